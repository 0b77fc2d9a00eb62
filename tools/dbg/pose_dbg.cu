// debug: run one projection step pieces on device and host, dump intermediates
#include <cstdio>
#include <cuda_runtime.h>
#include "../../mjpl_b200/csrc/vk_build.h"
using namespace vk;
struct Dump { double dx[6]; double J[6][8]; double A[6][6]; double y[6]; double site[7]; double anchor[8][3]; double axis[8][3]; };
VK_HD void one_step(const FkTables<double> &fk, int nslot, const PoseSpec &spec, const double *q, Dump &D) {
  Pose<double> P[MAX_BODY]; V3<double> anchor[MAX_JNT], axis[MAX_JNT];
  fk_with_joints(fk, nslot, q, P, anchor, axis);
  Pose<double> site; pose_displacement(spec, P, D.dx, site);
  D.site[0]=site.p.x; D.site[1]=site.p.y; D.site[2]=site.p.z; D.site[3]=site.q.w; D.site[4]=site.q.x; D.site[5]=site.q.y; D.site[6]=site.q.z;
  for (int j=0;j<fk.njnt&&j<8;j++){ D.anchor[j][0]=anchor[j].x; D.anchor[j][1]=anchor[j].y; D.anchor[j][2]=anchor[j].z; D.axis[j][0]=axis[j].x; D.axis[j][1]=axis[j].y; D.axis[j][2]=axis[j].z; }
  double roll,pitch,yaw; quat2rpy(site.q,roll,pitch,yaw);
  const double cp=cos(pitch), sp_=sin(pitch), cy=cos(yaw), sy=sin(yaw);
  const double E[3][3]={{cy/cp,sy/cp,0.0},{-sy,cp,0.0},{cy*(sp_/cp),sy*(sp_/cp),1.0}};
  double J[6][MAX_JNT];
  for (int j=0;j<fk.njnt;j++){ V3<double> jp=mk<double>(0,0,0), jr=mk<double>(0,0,0);
    if ((spec.jnt_mask>>j)&1u){ if (fk.jnt_type[j]==JK_SLIDE) jp=axis[j]; else { jr=axis[j]; jp=cross(axis[j], site.p-anchor[j]); } }
    const int c=fk.jnt_qadr[j];
    J[0][c]=jp.x;J[1][c]=jp.y;J[2][c]=jp.z;
    J[3][c]=E[0][0]*jr.x+E[0][1]*jr.y+E[0][2]*jr.z; J[4][c]=E[1][0]*jr.x+E[1][1]*jr.y+E[1][2]*jr.z; J[5][c]=E[2][0]*jr.x+E[2][1]*jr.y+E[2][2]*jr.z; }
  for (int i=0;i<6;i++) for (int c=0;c<8;c++) D.J[i][c]= c<fk.nq? J[i][c]:0;
  double A[6][6];
  for (int i=0;i<6;i++) for (int k=i;k<6;k++){ double acc=0; for (int c=0;c<fk.nq;c++) acc+=J[i][c]*J[k][c]; A[i][k]=A[k][i]=acc; }
  for (int i=0;i<6;i++) for (int k=0;k<6;k++) D.A[i][k]=A[i][k];
  sym6_pinv_apply(A, D.dx, D.y);
}
__global__ void k(const FkTables<double>* fk, int nslot, PoseSpec spec, const double* q, Dump* D){ one_step(*fk,nslot,spec,q,*D); }
extern "C" int run(const mjb_model_desc* d, const mjb_pose_spec* in, const double* q, Dump* host, Dump* dev){
  vkb::HostModel H; if(!vkb::build_host_model(d,H)) return 1;
  PoseSpec sp; std::string e; if(!vkb::make_pose_spec(H,in,sp,e)) return 2;
  one_step(H.fk,H.nslot,sp,q,*host);
  FkTables<double>* dfk; double* dq; Dump* dD;
  cudaMalloc(&dfk,sizeof(H.fk)); cudaMemcpy(dfk,&H.fk,sizeof(H.fk),cudaMemcpyHostToDevice);
  cudaMalloc(&dq,64*8); cudaMemcpy(dq,q,H.nq*8,cudaMemcpyHostToDevice);
  cudaMalloc(&dD,sizeof(Dump));
  k<<<1,1>>>(dfk,H.nslot,sp,dq,dD); cudaError_t er=cudaDeviceSynchronize(); if(er) {printf("cuda %s\n",cudaGetErrorString(er)); return 3;}
  cudaMemcpy(dev,dD,sizeof(Dump),cudaMemcpyDeviceToHost); return 0;
}
