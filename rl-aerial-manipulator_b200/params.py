"""Model constants of the reference quadrotor, computed with the same NumPy calls as
`simul_files/model/params.py:10-43` of the reference so that `invA` / `invI` are bit-identical.
The CUDA kernels receive these through `qs_config` (include/quadsim.h); nothing is hard-coded on the device.
"""
import numpy as np

mass = 0.18  # kg
g = 9.81  # m/s^2
I = np.array([(0.00025, 0, 2.55e-6), (0, 0.000232, 0), (2.55e-6, 0, 0.0003738)])
invI = np.linalg.inv(I)
arm_length = 0.086  # m
height = 0.05
minF = 0.0
maxF = 2.0 * mass * g
L = arm_length
H = height
km = 1.5e-9
kf = 6.11e-8
r = km / kf
A = np.array([[1, 1, 1, 1], [0, L, 0, -L], [-L, 0, L, 0], [r, -r, r, -r]])
invA = np.linalg.inv(A)
body_frame = np.array([(L, 0, 0, 1), (0, L, 0, 1), (-L, 0, 0, 1), (0, -L, 0, 1), (0, 0, 0, 1), (0, 0, H, 1)])
dt = 1.0 / 200.0  # WaypointQuadEnv.dt (rl_env_scaledObs.py:30)

ODEINT_TOL = 1.49012e-8  # scipy.integrate.odeint default rtol == atol
