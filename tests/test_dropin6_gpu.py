"""GPU: the reference-shaped single-vehicle surface of dynamicsModel_BlueROV2_Heavy_6DoF
(same class / method names and call conventions as the reference, numpy in / numpy out)
against golden vectors produced by the unmodified reference and SURVEY.md's known answers.
These read like the tests the reference never had."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from marinevehiclereinforcementlearning_b200 import dynamicsModel_BlueROV2_Heavy_6DoF as m6
    from marinevehiclereinforcementlearning_b200 import resources

KAT_STATE = np.array([0.1, -0.2, 0.3, 0.2, -0.1, 1.0, 0.3, -0.1, 0.05, 0.02, -0.03, 0.1])
KAT_RPM = np.array([1000., -2000., 500., 3600., -250., 1500., -1500., 800.])


def rel_err(a, ref):
    ref = np.asarray(ref)
    scale = np.abs(ref) + np.abs(ref).max(axis=-1, keepdims=True)
    return (np.abs(np.asarray(a) - ref) / np.maximum(scale, 1e-300)).max()


class ConstController:
    """The reference's controller seam: any object with these three members can be injected (6DoF.py:76-78, 418)."""
    def __init__(self, f):
        self.f, self.setPoint = np.asarray(f, dtype=float), np.zeros(6)

    def reset(self):
        pass

    def computeControlForces(self, x, y, z, phi, theta, psi, t):
        return self.f


def test_constants_and_allocation_kat():
    rov = m6.BlueROV2Heavy6DoF(ConstController(np.zeros(6)))
    assert rov.m == 11.4 and rov.Length == 0.457 and rov.CG[2] == 0.05 and rov.Zvdot == 0.0 and rov.Zwdot == -14.57
    assert rov.rho_f * rov.D_thruster ** 4 * rov.Kt_thruster == pytest.approx(0.011755102040816326, rel=1e-15)   # KAT-A
    assert np.allclose(rov.A[:, 0], [0.838671, -0.544639, 0, 0.037035, 0.05703, -0.16504], atol=1e-6)
    assert np.allclose(rov.Ainv[4], [-0.141667, -0.077273, -0.25, -1.136364, 2.083333, 0], atol=1e-6)
    r = load_golden("resources")
    A, Ainv = resources.computeThrustAllocation(rov.thrusterPositions, rov.thrusterNormals)
    assert np.array_equal(A, r["A6"]) and np.abs(Ainv - r["Ainv6"]).max() < 1e-15
    A0, A0inv = resources.computeThrustAllocation(rov.thrusterPositions, rov.thrusterNormals, x0=r["alloc_x0"])
    assert np.abs(A0 - r["A6_x0"]).max() < 1e-15 and np.abs(A0inv - r["Ainv6_x0"]).max() < 1e-13


def test_force_model_and_derivs_kat1():
    g = load_golden("rov6")
    rov = m6.BlueROV2Heavy6DoF(ConstController(np.zeros(6)))
    M, RHS = rov.forceModel(KAT_STATE[:3], KAT_STATE[3:6], KAT_STATE[6:], KAT_RPM)
    assert np.array_equal(M, g["M"])
    want = [-45.495385173564316, 12.808172228914902, -2.544429918367347, -4.700912058476027, -4.211373312264835, -8.967903250489707]
    assert rel_err(RHS, want) < 1e-13
    acc = np.linalg.solve(M, RHS)
    assert rel_err(acc, [-2.3458113182380824, 0.14117300847952502, -0.22319560687432868, -16.501583727295348,
                         -10.265217360246886, -32.028225894606095]) < 1e-12
    for i in (0, 17, 101):   # forceModel(retComp=True) -> 6 x 5 matrix of -Crb v, -Ca v, -D v, G, H
        comp = rov.forceModel(g["rpm_states"][i, :3], g["rpm_states"][i, 3:6], g["rpm_states"][i, 6:], g["rpm_rpms"][i], retComp=True)
        assert comp.shape == (6, 5) and np.abs(comp - g["rpm_retComp"][i]).max() < 1e-10
    assert np.abs(rov.thrusterModel(g["thruster_rpm"]) - g["thruster_F"]).max() < 1e-12
    assert rov.thrusterModel(3500.) == pytest.approx(40.0, rel=1e-14)


def test_moving_coordinate_system():
    g = load_golden("rov6")
    rov = m6.BlueROV2Heavy6DoF(ConstController(np.zeros(6)))
    for i in (0, 5, 200):
        rov.updateMovingCoordSystem(g["rpm_states"][i, 3:6])
        axes = np.array([rov.iHat, rov.jHat, rov.kHat])
        assert np.abs(axes - g["rpm_axes"][i]).max() < 1e-14
        v = np.array([0.3, -1.2, 2.0])
        assert np.abs(rov.globalToVehicle(v) - axes @ v).max() < 1e-14
        assert np.abs(rov.vehicleToGlobal(rov.globalToVehicle(v)) - v).max() < 1e-14


def test_derivs_with_injected_and_pid_controllers():
    g = load_golden("rov6")
    for i in (0, 3, 77):     # stateless injected controller == the reference's force mode
        rov = m6.BlueROV2Heavy6DoF(ConstController(g["force_forces"][i]))
        d = rov.derivs(0.0, g["force_states"][i])
        assert rel_err(d, g["force_derivs"][i]) < 1e-10
        assert np.abs(rov.controlVector - g["force_cv"][i]).max() < 1e-8
        assert np.abs(rov.allocateThrust() - g["force_cv"][i]).max() < 1e-8
    # KAT-2: the real PID, first call, set-point 0, t = 0.1
    pid = m6.BlueROV2Heavy6DoF_PID_controller(np.zeros(6))
    rov = m6.BlueROV2Heavy6DoF(pid)
    d = rov.derivs(0.1, KAT_STATE)
    assert np.abs(rov.generalisedControlForces - [-2.52, 5.04, -7.56, -1, 1, -1.02]).max() < 1e-12
    assert rel_err(d[6:], [-0.37836252889058264, 0.12903386553378404, -0.7717309882328138, -3.214521163590409,
                           8.018270531747259, -3.309355392109698]) < 1e-10
    assert abs(rov.controlVector[2] - (-1123.7517320894956)) < 1e-7
    # a PID call sequence in RK4 stage order (controller state mutates on every call, 6DoF.py:62-71)
    e = 2
    pid = m6.BlueROV2Heavy6DoF_PID_controller(g["pid_sp"][e].copy())
    rov = m6.BlueROV2Heavy6DoF(pid)
    for c in range(12):
        d = rov.derivs(g["pid_t"][e, c], g["pid_states"][e, c])
        assert rel_err(d, g["pid_derivs"][e, c]) < 1e-9, c
        assert np.abs(rov.generalisedControlForces - g["pid_gcf"][e, c]).max() < 1e-9
        assert np.abs(pid.eInt - g["pid_eint"][e, c]).max() < 1e-12
    pid.reset()
    assert pid.eOld is None and pid.tOld == 0.0


def test_gym_env_fixed_setpoint_episode_vs_reference():
    e6 = load_golden("env6")
    env = m6.BlueROV2Heavy6DoFEnv(maxSteps=60)
    assert env.action_space.shape == (6,) and env.observation_space.shape == (9,) and env.lenObs == 9
    obs = [env.reset(initialSetpoint=list(e6["fixed_sp"]))]
    assert env.fixedSp and env.iWp == 0 and np.array_equal(env.path[0], env.path[1])
    for k in range(60):
        ob, reward, done, info = env.step(np.zeros(6))
        obs.append(ob)
        assert reward == 0.0 and info == {} and done == bool(e6["fixed_done"][k])
    assert np.abs(np.array(obs) - e6["fixed_obs"]).max() < 1e-8
    h = env.timeHistory      # pandas DataFrame on done, the reference's 33 columns (6DoF.py:578-587)
    assert list(h.columns) == list(e6["fixed_history_cols"]) and h.shape == (61, 33)
    assert np.abs(h.values[:, :13] - e6["fixed_history"][:, :13]).max() < 1e-8
    assert np.abs(h.values[:, 27:] - e6["fixed_history"][:, 27:]).max() < 1e-12
    assert np.abs(env.dataToState(env.systemState) - env.state).max() < 1e-14
    # the reference's random branch raises (6DoF.py:497); here it draws a path
    env2 = m6.BlueROV2Heavy6DoFEnv(seed=3, maxSteps=5)
    ob = env2.reset()
    assert not env2.fixedSp and env2.path.shape == (2, 3) and np.abs(env2.path).max() <= 5.0 and ob.shape == (9,)
