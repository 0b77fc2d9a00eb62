for flags in "-O3" "-O3 -fmad=false" "-O1" "-O0" "-G"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 $flags -o /tmp/pinv_dbg tools/dbg/pinv_dbg.cu 2>&1 | grep -i error; echo "== $flags"; /tmp/pinv_dbg
done
