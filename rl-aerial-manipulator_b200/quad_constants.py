"""Physical constants of the simulated quadrotor and the matrices derived from them.

The values are the reference's (`simul_files/model/params.py:10-43`); the two inverses are taken with the same `np.linalg.inv`
call on the same float64 arrays, so `inv_mixer` / `inv_inertia` are bit-identical to the reference's `invA` / `invI` -- the
float64 parity mode depends on that.  The CUDA kernels receive all of this through `qs_config` (include/quadsim.h); nothing is
hard-coded on the device.  `QUAD` is the one instance; `_cabi.make_config` serialises it.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def _inertia() -> np.ndarray:
    ixx, iyy, izz, ixz = 2.5e-4, 2.32e-4, 3.738e-4, 2.55e-6          # kg m^2, body axes
    return np.array([(ixx, 0, ixz), (0, iyy, 0), (ixz, 0, izz)])


def _mixer(arm: float, torque_per_thrust: float) -> np.ndarray:
    """(F, M1, M2, M3) = mixer @ (thrust of rotors 1..4): plus-configuration arms, alternating spin directions."""
    a, r = arm, torque_per_thrust
    return np.array([[1, 1, 1, 1], [0, a, 0, -a], [-a, 0, a, 0], [r, -r, r, -r]])


@dataclass(frozen=True)
class QuadrotorConstants:
    mass: float = 0.18                      # kg
    gravity: float = 9.81                   # m/s^2
    arm_length: float = 0.086               # m, hub to rotor
    body_height: float = 0.05               # m, drawn mast of the plotting frame
    rotor_drag_coeff: float = 1.5e-9        # km
    rotor_thrust_coeff: float = 6.11e-8     # kf
    min_total_thrust: float = 0.0           # N
    control_dt: float = 1.0 / 200.0         # s, WaypointQuadEnv.dt (rl_env_scaledObs.py:30)
    odeint_tol: float = 1.49012e-8          # scipy.integrate.odeint default rtol == atol
    inertia: np.ndarray = field(default_factory=_inertia)

    @property
    def max_total_thrust(self) -> float:    # twice the hover thrust
        return 2.0 * self.mass * self.gravity

    @property
    def torque_per_thrust(self) -> float:
        return self.rotor_drag_coeff / self.rotor_thrust_coeff

    @property
    def mixer(self) -> np.ndarray:
        return _mixer(self.arm_length, self.torque_per_thrust)

    @property
    def inv_mixer(self) -> np.ndarray:
        return np.linalg.inv(self.mixer)

    @property
    def inv_inertia(self) -> np.ndarray:
        return np.linalg.inv(self.inertia)

    @property
    def body_frame(self) -> np.ndarray:
        """Homogeneous body-frame points the plotting scripts draw: four rotor hubs, the centre, the mast top."""
        a, h = self.arm_length, self.body_height
        return np.array([(a, 0, 0, 1), (0, a, 0, 1), (-a, 0, 0, 1), (0, -a, 0, 1), (0, 0, 0, 1), (0, 0, h, 1)])


QUAD = QuadrotorConstants()
