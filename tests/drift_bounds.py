"""Stated multi-step drift bounds of the fixed-step modes against the reference's own trajectories.

The reference integrates with adaptive LSODA (rtol = atol = 1.49e-8 per internal step), so a fixed-step RK4 trajectory under
the same actions separates from it at the rate of LSODA's own truncation error; in float32 the rounding of the state adds to
that.  Bound after h env steps since the two trajectories were last identical (episode start):

        |x_ours - x_reference|  <=  C * (1 + h / 50)

with C per quantity and mode below.  Two action regimes are pinned:
  "traj"    tests/golden/traj_v{1,2}.npz -- hover, uniform-random and P-controller actions, 600 steps with auto-reset
  "policy"  tests/golden/closed_loop_v2.npz -- the action sequences the shipped v2 policy (ppo_model_2300000_steps.zip) produced
            on the reference env (bang-bang: most actions sit on the box limits), whole episodes of 630-840 steps
C was set at ~4x the worst value measured with the device code compiled for the host (tests/test_host_harness.py) and holds on
the B200 (tests/test_gpu_trajectory.py).  Flags (terminated / truncated / info bits) must be equal at every step.
"""

# quantity -> C; pos [m], vel [m/s], quat [-], omega [rad/s], reward [-], obs [scaled units]
import numpy as np

BOUNDS = {
    # reward: the B200 build (FMA contraction, device intrinsics) measured 1.163e-3 at h = 529, i.e. C = 1.005e-4, against 2.5e-5 for
    # the host build of the same code -- the stated C keeps 1.5x over the worst device value
    ("traj", "f32"): dict(pos=2e-5, vel=3e-5, quat=2e-6, omega=1e-5, reward=1.5e-4, obs=1e-5),
    ("traj", "f64"): dict(pos=2e-6, vel=2e-6, quat=1e-7, omega=2e-6, reward=2e-5, obs=1e-6),
    ("policy", "f32"): dict(pos=1e-4, vel=1e-4, quat=1e-5, omega=1e-5, reward=2e-3, obs=5e-5),
    ("policy", "f64"): dict(pos=3e-5, vel=3e-5, quat=5e-6, omega=5e-6, reward=5e-4, obs=2e-5),
}
# the +2 progress bonus of _calculate_reward flips when |last_distance - distance| is below the resolution of the mode
BONUS_FLIP_WINDOW = {"f32": 5e-6, "f64": 2e-7}


def bound(regime: str, precision: str, quantity: str, h):
    return BOUNDS[(regime, precision)][quantity] * (1.0 + h / 50.0)


def check_step(regime, precision, h, y, y_ref, reward, reward_ref, dd_ref, obs=None, obs_ref=None, tag=""):
    """state / reward / obs of the envs in this step against the reference, each env at its own horizon h."""
    e = np.abs(y - y_ref)
    for name, sl in (("pos", slice(0, 3)), ("vel", slice(3, 6)), ("quat", slice(6, 10)), ("omega", slice(10, 13))):
        lim = bound(regime, precision, name, h)
        assert np.all(e[:, sl].max(1) <= lim), f"{tag} {name}: {e[:, sl].max(1)} > {lim} at h={h}"
    er = np.abs(reward - reward_ref)
    lim = bound(regime, precision, "reward", h)
    # the +2 progress bonus may flip when the reference's own change of distance is below the mode's resolution
    flip = (np.abs(er - 2.0) <= lim) & (np.abs(dd_ref) < BONUS_FLIP_WINDOW[precision] * (1.0 + h / 50.0))
    assert np.all((er <= lim) | flip), f"{tag} reward: {er} > {lim} at h={h}"
    if obs is not None:
        eo = np.abs(obs.astype(np.float64) - obs_ref.astype(np.float64)).max(1)
        assert np.all(eo <= bound(regime, precision, "obs", h)), f"{tag} obs: {eo} at h={h}"
    return int(flip.sum())
