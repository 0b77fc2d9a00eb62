"""Small MJCF models written for these tests (same geometry as the scenes the reference's unit
tests describe: a ball sliding along x / in the xy-plane next to a wall)."""

ONE_DOF_BALL = """
<mujoco model="ball on a rail">
  <worldbody>
    <geom type="plane" size="2 2 0.1"/>
    <body name="ball" pos="0 0 1">
      <joint name="ball_slide_x" type="slide" axis="1 0 0" range="-2 2"/>
      <geom type="sphere" size="0.01"/>
    </body>
    <geom name="wall_obstacle" type="box" pos="0.9 0 1" size="0.05 0.5 0.5"/>
  </worldbody>
</mujoco>
"""

TWO_DOF_BALL = """
<mujoco model="ball in a plane">
  <worldbody>
    <geom type="plane" size="2 2 0.1"/>
    <body name="ball" pos="0 0 1">
      <joint name="ball_slide_x" type="slide" axis="1 0 0" range="-2 2"/>
      <joint name="ball_slide_y" type="slide" axis="0 1 0" range="-2 2"/>
      <geom type="sphere" size="0.1"/>
      <site name="ball_site"/>
    </body>
    <geom name="wall_obstacle" type="box" pos="0.6 0 1" size="0.1 0.5 0.5"/>
  </worldbody>
</mujoco>
"""

JOINT_ZOO = """
<mujoco model="joint zoo">
  <worldbody>
    <body><geom size="1"/><joint name="slide_joint" type="slide"/></body>
    <body><geom size="1"/><freejoint name="free_joint"/>
      <body><geom size="1"/><joint name="hinge_joint" type="hinge"/></body>
    </body>
    <body><geom size="1"/><joint name="ball_joint" type="ball"/></body>
  </worldbody>
</mujoco>
"""

# every primitive pair type, nested defaults, childclass, fromto, excludes, euler/degree angles
PRIMITIVE_ARM = """
<mujoco model="primitive arm">
  <compiler angle="degree" autolimits="true"/>
  <default>
    <default class="arm">
      <joint axis="0 0 1" range="-170 170"/>
      <geom type="capsule" size="0.04"/>
      <default class="tip"><geom type="box" size="0.03 0.02 0.05"/></default>
      <default class="ghost"><geom contype="0" conaffinity="0"/></default>
    </default>
  </default>
  <worldbody>
    <geom name="floor" type="plane" size="0 0 0.05"/>
    <geom name="pillar" type="cylinder" size="0.05 0.3" pos="0.45 0.1 0.3"/>
    <geom name="ball" type="sphere" size="0.08" pos="-0.3 0.35 0.5"/>
    <geom name="beam" type="capsule" size="0.03" fromto="-0.5 -0.4 0.7 0.5 -0.4 0.7"/>
    <geom name="crate" type="box" size="0.1 0.1 0.1" pos="0.0 0.55 0.1" euler="0 0 30"/>
    <body name="base" pos="0 0 0.1" childclass="arm">
      <geom name="base_geom" type="cylinder" size="0.08 0.1"/>
      <body name="l1" pos="0 0 0.15">
        <joint name="j1"/>
        <geom name="l1_geom" fromto="0 0 0 0 0 0.3"/>
        <geom class="ghost" type="sphere" size="0.5"/>
        <body name="l2" pos="0 0 0.3" euler="90 0 0">
          <joint name="j2" range="-120 120"/>
          <geom name="l2_geom" fromto="0 0 0 0.3 0 0"/>
          <body name="l3" pos="0.3 0 0">
            <joint name="j3" range="-150 150"/>
            <geom name="l3_geom" fromto="0 0 0 0.25 0 0" size="0.03"/>
            <geom name="l3_ball" type="sphere" size="0.05" pos="0.12 0.05 0"/>
            <body name="wrist" pos="0.25 0 0">
              <joint name="j4" axis="1 0 0"/>
              <geom name="tip_geom" class="tip" pos="0.05 0 0"/>
              <body name="finger" pos="0.1 0 0">
                <joint name="j5" type="slide" axis="0 1 0" range="0 0.04"/>
                <geom name="finger_geom" type="box" size="0.01 0.01 0.03"/>
              </body>
            </body>
          </body>
        </body>
      </body>
    </body>
  </worldbody>
  <contact><exclude body1="l1" body2="l3"/></contact>
  <keyframe><key name="home" qpos="0 0.5 -1.0 0 0.02"/></keyframe>
</mujoco>
"""
