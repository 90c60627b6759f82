"""Benchmark / parity workloads, written as *user scripts* against the public API.

Every builder takes an ``api`` namespace that provides ``Problem, Guess,
Condition, Dynamics`` -- the drop-in facade (``OpenGoddard.optimize``), the
numpy oracle (``oracle.og_numpy``) or, inside the build container, the real
reference module -- and returns a fully set-up problem (guess written, callbacks
assigned).  The same definition therefore drives the CUDA path, the oracle and
the reference, which is what makes the parity tests meaningful.

The optimal-control problems are the ones `BASELINE.json:configs` names; the
physics follows the reference's shipped example scripts (cited per builder) with
the same operation order, so constraint vectors agree with the examples to the
last bit, but the node counts are parameters.

Synthetic instance batches (`make_batch`) follow SURVEY.md section 8(d).
"""
from collections import namedtuple

import numpy as np

Workload = namedtuple("Workload", "name prob obj sanitize")

SEED0 = 20261017


# --------------------------------------------------------------------------
# cfg1: Brachistochrone (reference: examples/01_Brachistochrone_Problem.py:9-91)
# --------------------------------------------------------------------------
class _Bead:
    def __init__(self):
        self.g = 1.0
        self.l = 1.0


def brachistochrone(api, nodes=(20,)):
    bead = _Bead()
    prob = api.Problem([0.0, 2.0], list(nodes), [3], [1], 30)

    def dyn(prob, obj, section):
        v = prob.states(2, section)
        th = prob.controls(0, section)
        d = api.Dynamics(prob, section)
        d[0] = v * np.sin(th)
        d[1] = v * np.cos(th)
        d[2] = obj.g * np.cos(th)
        return d()

    def eq(prob, obj):
        x = prob.states_all_section(0)
        y = prob.states_all_section(1)
        v = prob.states_all_section(2)
        r = api.Condition()
        r.equal(x[0], 0.0)
        r.equal(y[0], 0.0)
        r.equal(v[0], 0.0)
        r.equal(x[-1], obj.l)
        return r()

    def ineq(prob, obj):
        y = prob.states_all_section(1)
        th = prob.controls_all_section(0)
        tf = prob.time_final(-1)
        r = api.Condition()
        r.lower_bound(tf, 0.1)
        r.lower_bound(y, 0)
        r.lower_bound(th, 0)
        return r()

    def cost(prob, obj):
        return prob.time_final(-1)

    def cost_grad(prob, obj):
        g = api.Condition(prob.number_of_variables)
        g.change_value(prob.index_time_final(-1), 1)
        return g()

    t = prob.time_all_section
    prob.set_states_all_section(0, api.Guess.linear(t, 0.0, bead.l))
    prob.set_states_all_section(1, api.Guess.linear(t, 0.0, bead.l / np.sqrt(3)))
    prob.set_controls_all_section(0, api.Guess.linear(t, np.deg2rad(30), np.deg2rad(30)))
    prob.dynamics = [dyn]
    prob.knot_states_smooth = []
    prob.cost = cost
    prob.cost_derivative = cost_grad
    prob.equality = eq
    prob.inequality = ineq
    return Workload("brachistochrone", prob, bead, None)


# --------------------------------------------------------------------------
# cfg2 / cfg3: Goddard rocket (reference: examples/04_Goddard_0knot.py:10-97,
# examples/05_Goddard_1knot.py:10-175)
# --------------------------------------------------------------------------
class _GoddardRocket:
    def __init__(self):
        self.g0 = 1.0
        self.H0, self.V0, self.M0 = 1.0, 0.0, 1.0
        Tc, Vc, Mc = 3.5, 620, 0.6
        self.Hc = 500
        self.c = 0.5 * np.sqrt(self.g0 * self.H0)
        self.Mf = Mc * self.M0
        self.Dc = 0.5 * Vc * self.M0 / self.g0
        self.T_max = Tc * self.g0 * self.M0


def _goddard_callbacks(api, two_phase):
    def dyn(prob, obj, section):
        h = prob.states(0, section)
        v = prob.states(1, section)
        m = prob.states(2, section)
        T = prob.controls(0, section)
        drag = 1 * obj.Dc * v ** 2 * np.exp(-obj.Hc * (h - obj.H0) / obj.H0)
        grav = obj.g0 * (obj.H0 / h) ** 2
        d = api.Dynamics(prob, section)
        d[0] = v
        d[1] = (T - drag) / m - grav
        d[2] = - T / obj.c
        return d()

    def eq(prob, obj):
        h = prob.states_all_section(0)
        v = prob.states_all_section(1)
        m = prob.states_all_section(2)
        r = api.Condition()
        r.equal(h[0], obj.H0)
        r.equal(v[0], obj.V0)
        r.equal(m[0], obj.M0)
        r.equal(v[-1], 0.0)
        r.equal(m[-1], obj.Mf)
        if two_phase:
            r.equal(prob.time_final(0), 0.075)
        return r()

    def ineq(prob, obj):
        h = prob.states_all_section(0)
        v = prob.states_all_section(1)
        m = prob.states_all_section(2)
        T = prob.controls_all_section(0)
        tf = prob.time_final(-1)
        r = api.Condition()
        r.lower_bound(h, obj.H0)
        r.lower_bound(v, 0.0)
        r.lower_bound(m, obj.Mf)
        r.lower_bound(T, 0.0)
        r.lower_bound(tf, 0.1)
        if two_phase:
            r.lower_bound(prob.time_final(0), 0.05)
        r.upper_bound(m, obj.M0)
        r.upper_bound(T, obj.T_max)
        return r()

    return dyn, eq, ineq


def goddard(api, nodes=(50,)):
    """Goddard 0-knot: 1 phase, 3 states (h, v, m), 1 control (thrust)."""
    ro = _GoddardRocket()
    prob = api.Problem([0.0, 0.3], list(nodes), [3], [1], 30)
    dyn, eq, ineq = _goddard_callbacks(api, two_phase=False)

    def cost(prob, obj):
        return -prob.states_all_section(0)[-1]

    t = prob.time_all_section
    prob.set_states_all_section(0, api.Guess.cubic(t, 1.0, 0.0, 1.010, 0.0))
    prob.set_states_all_section(1, api.Guess.linear(t, 0.0, 0.0))
    prob.set_states_all_section(2, api.Guess.cubic(t, 1.0, -0.6, 0.6, 0.0))
    prob.set_controls_all_section(0, api.Guess.cubic(t, 3.5, 0.0, 0.0, 0.0))
    prob.dynamics = [dyn]
    prob.knot_states_smooth = []
    prob.cost = cost
    prob.cost_derivative = None
    prob.equality = eq
    prob.inequality = ineq

    def sanitize(prob, P):
        n0 = prob.nodes[0]
        P[:, 0:n0] = np.maximum(P[:, 0:n0], 0.99)          # h >= 0.99
        P[:, 2 * n0:3 * n0] = np.maximum(P[:, 2 * n0:3 * n0], 0.1)  # m >= 0.1
        return P

    return Workload("goddard", prob, ro, sanitize)


def goddard_knot(api, nodes=(25, 25)):
    """Goddard 1-knot: 2 phases joined by smooth-state knot rows; state 0 carries a
    canonical unit of 0.1 and the cost reads ``states_all_section(-1)`` exactly as
    the shipped script does."""
    ro = _GoddardRocket()
    prob = api.Problem([0.0, 0.1, 0.3], list(nodes), [3, 3], [1, 1], 50)
    prob.set_unit_states_all_section(0, 0.1)
    dyn, eq, ineq = _goddard_callbacks(api, two_phase=True)

    def cost(prob, obj):
        return -prob.states_all_section(-1)[-1]

    def cost_grad(prob, obj):
        g = api.Condition(prob.number_of_variables)
        g.change_value(prob.index_states(0, -1, -1), -1)
        return g()

    t = prob.time_all_section
    prob.set_states_all_section(0, api.Guess.cubic(t, 1.0, 0.0, 1.010, 0.0))
    prob.set_states_all_section(1, api.Guess.linear(t, 0.0, 0.0))
    prob.set_states_all_section(2, np.hstack((api.Guess.linear(prob.time[0], 1.0, 0.6),
                                              api.Guess.linear(prob.time[1], 0.6, 0.6))))
    prob.set_controls_all_section(0, np.hstack((api.Guess.linear(prob.time[0], 3.5, 3.5),
                                                api.Guess.linear(prob.time[1], 0.0, 0.0))))
    prob.dynamics = [dyn, dyn]
    prob.knot_states_smooth = [True]
    prob.cost = cost
    prob.cost_derivative = cost_grad
    prob.equality = eq
    prob.inequality = ineq

    def sanitize(prob, P):
        u0 = prob.unit_states[0][0]
        for s in range(prob.number_of_section):
            a = prob.index_states(0, s)
            P[:, a:a + prob.nodes[s]] = np.maximum(P[:, a:a + prob.nodes[s]], 0.99 / u0)
            a = prob.index_states(2, s)
            P[:, a:a + prob.nodes[s]] = np.maximum(P[:, a:a + prob.nodes[s]], 0.1)
        return P

    return Workload("goddard_knot", prob, ro, sanitize)


# --------------------------------------------------------------------------
# cfg4: polar-coordinate multi-stage ascent
# (reference: examples/09_Rocket_Ascent_Polar_TSTO.py:10-293)
# --------------------------------------------------------------------------
class _StagedLauncher:
    GMe = 3.986004418 * 10 ** 14
    Re = 6371.0 * 1000
    g0 = 9.80665

    def __init__(self, nstage_sections):
        self.M0 = [20000.0, 1000.0]
        self.Mdry = [2000, 200]
        self.Cd = [0.2, 0.2]
        self.A = [3.14, 3.14]
        self.Isp = [300.0, 350.0]
        self.Tmax = [self.M0[0] * self.g0 * 1.5, self.M0[1] * self.g0 * 1.5]
        self.MaxG = 8.0
        self.Rtarget = self.Re + 500.0 * 1000
        self.Vtarget = np.sqrt(self.GMe / self.Rtarget)
        # section -> stage map (2-phase: [0, 1]; 3-phase extension: [0, 1, 1])
        self.stage = nstage_sections

    def air_density(self, h):
        beta = 1 / 8500.0
        rho0 = 1.225
        h[h < -100.0] = -100.0
        return rho0 * np.exp(-beta * h)


def _polar_dynamics(api):
    def dyn(prob, obj, section):
        R = prob.states(0, section)
        Vr = prob.states(2, section)
        Vt = prob.states(3, section)
        m = prob.states(4, section)
        Tr = prob.controls(0, section)
        Tt = prob.controls(1, section)
        st = obj.stage[section]
        rho = obj.air_density(R - obj.Re)
        Dr = 0.5 * rho * Vr * np.sqrt(Vr ** 2 + Vt ** 2) * obj.Cd[st] * obj.A[st]
        Dt = 0.5 * rho * Vt * np.sqrt(Vr ** 2 + Vt ** 2) * obj.Cd[st] * obj.A[st]
        grav = obj.g0 * (obj.Re / R) ** 2
        d = api.Dynamics(prob, section)
        d[0] = Vr
        d[1] = Vt / R
        d[2] = Tr / m - Dr / m - grav + Vt ** 2 / R
        d[3] = Tt / m - Dt / m - (Vr * Vt) / R
        d[4] = - np.sqrt(Tr ** 2 + Tt ** 2) / obj.g0 / obj.Isp[st]
        return d()
    return dyn


def _polar_setup_units(prob, veh):
    uR = veh.Re
    uV = np.sqrt(veh.GMe / veh.Re)
    um = veh.M0[0]
    ut = uR / uV
    uT = um * uR / ut ** 2
    for k, u in enumerate((uR, 1, uV, uV, um)):
        prob.set_unit_states_all_section(k, u)
    prob.set_unit_controls_all_section(0, uT)
    prob.set_unit_controls_all_section(1, uT)
    prob.set_unit_time(ut)


def polar_tsto(api, nodes=(20, 20)):
    """Two-stage polar ascent, 2 phases; knot linkage lives in the user equality
    (``knot_states_smooth=[False]``)."""
    veh = _StagedLauncher([0, 1])
    prob = api.Problem([0.0, 100, 200], list(nodes), [5, 5], [2, 2], 40)
    _polar_setup_units(prob, veh)
    dyn = _polar_dynamics(api)

    def eq(prob, obj):
        Vr = prob.states_all_section(2)
        Vt = prob.states_all_section(3)
        R0, R1 = prob.states(0, 0), prob.states(0, 1)
        th0, th1 = prob.states(1, 0), prob.states(1, 1)
        Vr0, Vr1 = prob.states(2, 0), prob.states(2, 1)
        Vt0, Vt1 = prob.states(3, 0), prob.states(3, 1)
        m0, m1 = prob.states(4, 0), prob.states(4, 1)
        uR, uV, um = prob.unit_states[0][0], prob.unit_states[0][2], prob.unit_states[0][4]
        r = api.Condition()
        r.equal(R0[0], obj.Re, unit=uR)
        r.equal(th0[0], 0.0)
        r.equal(Vr0[0], 0.0, unit=uV)
        r.equal(Vt0[0], 0.0, unit=uV)
        r.equal(m0[0], obj.M0[0], unit=um)
        r.equal(m1[0], obj.M0[1], unit=um)
        r.equal(R1[-1], obj.Rtarget, unit=uR)
        r.equal(Vr[-1], 0.0, unit=uV)
        r.equal(Vt[-1], obj.Vtarget, unit=uV)
        r.equal(R1[0], R0[-1], unit=uR)
        r.equal(th1[0], th0[-1])
        r.equal(Vr1[0], Vr0[-1], unit=uV)
        r.equal(Vt1[0], Vt0[-1], unit=uV)
        return r()

    def ineq(prob, obj):
        R = prob.states_all_section(0)
        Vr = prob.states_all_section(2)
        Vt = prob.states_all_section(3)
        m = prob.states_all_section(4)
        Tr = prob.controls_all_section(0)
        Tt = prob.controls_all_section(1)
        Tr0, Tr1 = prob.controls(0, 0), prob.controls(0, 1)
        Tt0, Tt1 = prob.controls(1, 0), prob.controls(1, 1)
        rho = obj.air_density(R - obj.Re)
        Dr0 = 0.5 * rho * Vr * np.sqrt(Vr ** 2 + Vt ** 2) * obj.Cd[0] * obj.A[0]
        Dt0 = 0.5 * rho * Vt * np.sqrt(Vr ** 2 + Vt ** 2) * obj.Cd[0] * obj.A[0]
        Dr1 = 0.5 * rho * Vr * np.sqrt(Vr ** 2 + Vt ** 2) * obj.Cd[1] * obj.A[1]
        Dt1 = 0.5 * rho * Vt * np.sqrt(Vr ** 2 + Vt ** 2) * obj.Cd[1] * obj.A[1]
        a_r0 = (Tr - Dr0) / m
        a_t0 = (Tt - Dt0) / m
        a_mag0 = np.sqrt(a_r0 ** 2 + a_t0 ** 2)
        a_r1 = (Tr - Dr1) / m
        a_t1 = (Tt - Dt1) / m
        a_mag1 = np.sqrt(a_r1 ** 2 + a_t1 ** 2)
        T0 = np.sqrt(Tr0 ** 2 + Tt0 ** 2)
        T1 = np.sqrt(Tr1 ** 2 + Tt1 ** 2)
        r = api.Condition()
        r.lower_bound(R, obj.Re, unit=prob.unit_states[0][0])
        r.upper_bound(T0, obj.Tmax[0], unit=prob.unit_controls[0][0])
        r.upper_bound(T1, obj.Tmax[1], unit=prob.unit_controls[0][0])
        r.upper_bound(a_mag0, obj.MaxG * obj.g0)
        r.upper_bound(a_mag1, obj.MaxG * obj.g0)
        return r()

    def cost(prob, obj):
        return -prob.states(4, 1)[-1] / prob.unit_states[1][4]

    t = prob.time_all_section
    G = api.Guess
    prob.set_states_all_section(0, G.cubic(t, veh.Re, 0.0, veh.Rtarget, 0.0))
    prob.set_states_all_section(1, G.cubic(t, 0.0, 0.0, np.deg2rad(25.0), 0.0))
    prob.set_states_all_section(2, G.linear(t, 0.0, 0.0))
    prob.set_states_all_section(3, G.linear(t, 0.0, veh.Vtarget))
    # the shipped script hands a 2x-long mass guess; only the first sum(nodes) are used
    prob.set_states_all_section(4, np.hstack((G.cubic(t, veh.M0[0], -0.6, veh.Mdry[0], 0.0),
                                              G.cubic(t, veh.M0[1], -0.6, veh.Mdry[1], 0.0))))
    Tr_guess = np.hstack((G.cubic(prob.time[0], veh.Tmax[0] * 9 / 10, 0.0, 0.0, 0.0),
                          G.cubic(prob.time[1], veh.Tmax[1] * 9 / 10, 0.0, 0.0, 0.0)))
    prob.set_controls_all_section(0, Tr_guess)
    prob.set_controls_all_section(1, Tr_guess)   # (sic) the shipped script reuses Tr here
    prob.set_states_bounds_all_section(0, veh.Re, None)
    prob.set_states_bounds(4, 0, veh.Mdry[0], veh.M0[0])
    prob.set_states_bounds(4, 1, 1.0, veh.M0[1])
    prob.set_controls_bounds(0, 0, -veh.Tmax[1], veh.Tmax[0])
    prob.set_controls_bounds(1, 0, -veh.Tmax[1], veh.Tmax[0])
    prob.set_controls_bounds(0, 1, -veh.Tmax[1], veh.Tmax[1])
    prob.set_controls_bounds(1, 1, -veh.Tmax[1], veh.Tmax[1])
    prob.dynamics = [dyn, dyn]
    prob.knot_states_smooth = [False]
    prob.cost = cost
    prob.equality = eq
    prob.inequality = ineq
    return Workload("polar_tsto", prob, veh, None)


def polar_3phase(api, nodes=(40, 40, 40)):
    """cfg4: a *synthetic* 3-phase extension of the two-stage polar ascent (the
    shipped example has 2 phases; SURVEY.md section 2.3 note ii): stage 1 burn,
    stage 2 first burn, stage 2 second burn.  Knot 0 (staging, mass reset) is
    linked in the user equality; knot 1 uses the built-in smooth-state knot rows
    (``knot_states_smooth=[False, True]``)."""
    veh = _StagedLauncher([0, 1, 1])
    prob = api.Problem([0.0, 100, 160, 220], list(nodes), [5, 5, 5], [2, 2, 2], 40)
    _polar_setup_units(prob, veh)
    dyn = _polar_dynamics(api)

    def eq(prob, obj):
        uR, uV, um = prob.unit_states[0][0], prob.unit_states[0][2], prob.unit_states[0][4]
        first = [prob.states(k, 0) for k in range(5)]
        second = [prob.states(k, 1) for k in range(5)]
        last = [prob.states(k, 2) for k in range(5)]
        r = api.Condition()
        r.equal(first[0][0], obj.Re, unit=uR)
        r.equal(first[1][0], 0.0)
        r.equal(first[2][0], 0.0, unit=uV)
        r.equal(first[3][0], 0.0, unit=uV)
        r.equal(first[4][0], obj.M0[0], unit=um)
        r.equal(second[4][0], obj.M0[1], unit=um)
        r.equal(last[0][-1], obj.Rtarget, unit=uR)
        r.equal(last[2][-1], 0.0, unit=uV)
        r.equal(last[3][-1], obj.Vtarget, unit=uV)
        r.equal(second[0][0], first[0][-1], unit=uR)
        r.equal(second[1][0], first[1][-1])
        r.equal(second[2][0], first[2][-1], unit=uV)
        r.equal(second[3][0], first[3][-1], unit=uV)
        return r()

    def ineq(prob, obj):
        R = prob.states_all_section(0)
        Vr = prob.states_all_section(2)
        Vt = prob.states_all_section(3)
        m = prob.states_all_section(4)
        Tr = prob.controls_all_section(0)
        Tt = prob.controls_all_section(1)
        rho = obj.air_density(R - obj.Re)
        speed = np.sqrt(Vr ** 2 + Vt ** 2)
        Dr = 0.5 * rho * Vr * speed * obj.Cd[0] * obj.A[0]
        Dt = 0.5 * rho * Vt * speed * obj.Cd[0] * obj.A[0]
        a_mag = np.sqrt(((Tr - Dr) / m) ** 2 + ((Tt - Dt) / m) ** 2)
        r = api.Condition()
        r.lower_bound(R, obj.Re, unit=prob.unit_states[0][0])
        for s in range(3):
            Ts = np.sqrt(prob.controls(0, s) ** 2 + prob.controls(1, s) ** 2)
            r.upper_bound(Ts, obj.Tmax[obj.stage[s]], unit=prob.unit_controls[0][0])
        r.upper_bound(a_mag, obj.MaxG * obj.g0)
        r.lower_bound(prob.time_final(0), 30.0, unit=prob.unit_time)
        return r()

    def cost(prob, obj):
        return -prob.states(4, 2)[-1] / prob.unit_states[2][4]

    t = prob.time_all_section
    G = api.Guess
    n0, n1, n2 = nodes
    prob.set_states_all_section(0, G.cubic(t, veh.Re, 0.0, veh.Rtarget, 0.0))
    prob.set_states_all_section(1, G.cubic(t, 0.0, 0.0, np.deg2rad(25.0), 0.0))
    prob.set_states_all_section(2, G.linear(t, 0.0, 0.0))
    prob.set_states_all_section(3, G.linear(t, 0.0, veh.Vtarget))
    t12 = np.concatenate((prob.time[1], prob.time[2]))
    prob.set_states_all_section(4, np.hstack((G.linear(prob.time[0], veh.M0[0], veh.Mdry[0]),
                                              G.linear(t12, veh.M0[1], veh.Mdry[1]))))
    Tg = np.hstack((G.cubic(prob.time[0], veh.Tmax[0] * 9 / 10, 0.0, 0.0, 0.0),
                    G.cubic(t12, veh.Tmax[1] * 9 / 10, 0.0, 0.0, 0.0)))
    prob.set_controls_all_section(0, Tg)
    prob.set_controls_all_section(1, Tg * 0.25)
    prob.set_states_bounds_all_section(0, veh.Re, None)
    prob.set_states_bounds(4, 0, veh.Mdry[0], veh.M0[0])
    prob.set_states_bounds(4, 1, 1.0, veh.M0[1])
    prob.set_states_bounds(4, 2, 1.0, veh.M0[1])
    for s in range(3):
        hi = veh.Tmax[veh.stage[s]]
        prob.set_controls_bounds(0, s, -veh.Tmax[1], hi)
        prob.set_controls_bounds(1, s, -veh.Tmax[1], hi)
    prob.dynamics = [dyn, dyn, dyn]
    prob.knot_states_smooth = [False, True]
    prob.cost = cost
    prob.equality = eq
    prob.inequality = ineq
    return Workload("polar_3phase", prob, veh, None)


# --------------------------------------------------------------------------
# cfg5: low-thrust orbit transfer with a running (integrated) cost
# (reference: examples/10_Low_Thrust_Orbit_Transfer.py:10-168)
# --------------------------------------------------------------------------
class _Spacecraft:
    def __init__(self):
        self.u_max = 0.01
        self.r0, self.vr0, self.vt0 = 1.0, 0.0, 1.0
        self.rf, self.vrf, self.vtf = 4.0, 0.0, 0.5
        self.tf_max = 55


def low_thrust(api, nodes=(100,)):
    sc = _Spacecraft()
    prob = api.Problem([0.0, 10.0], list(nodes), [3], [4], 10)

    def dyn(prob, obj, section):
        r = prob.states(0, section)
        vr = prob.states(1, section)
        vt = prob.states(2, section)
        ur1 = prob.controls(0, section)
        ur2 = prob.controls(1, section)
        ut1 = prob.controls(2, section)
        ut2 = prob.controls(3, section)
        d = api.Dynamics(prob, section)
        d[0] = vr
        d[1] = vt ** 2 / r - 1 / r ** 2 + (ur1 - ur2)
        d[2] = - vr * vt / r + (ut1 - ut2)
        return d()

    def eq(prob, obj):
        r = prob.states_all_section(0)
        vr = prob.states_all_section(1)
        vt = prob.states_all_section(2)
        c = api.Condition()
        c.equal(r[0], obj.r0)
        c.equal(vr[0], obj.vr0)
        c.equal(vt[0], obj.vt0)
        c.equal(r[-1], obj.rf)
        c.equal(vr[-1], obj.vrf)
        c.equal(vt[-1], obj.vtf)
        return c()

    def ineq(prob, obj):
        r = prob.states_all_section(0)
        ur1 = prob.controls_all_section(0)
        ur2 = prob.controls_all_section(1)
        ut1 = prob.controls_all_section(2)
        ut2 = prob.controls_all_section(3)
        tf = prob.time_final(-1)
        c = api.Condition()
        c.lower_bound(r, obj.r0)
        c.lower_bound(ur1, 0.0)
        c.lower_bound(ut1, 0.0)
        c.lower_bound(ur2, 0.0)
        c.lower_bound(ut2, 0.0)
        c.lower_bound(tf, 0.0)
        c.upper_bound(r, obj.rf)
        c.upper_bound(ur1, obj.u_max)
        c.upper_bound(ut1, obj.u_max)
        c.upper_bound(ur2, obj.u_max)
        c.upper_bound(ut2, obj.u_max)
        c.upper_bound(tf, obj.tf_max)
        return c()

    def cost(prob, obj):
        return 0.0

    def running(prob, obj):
        ur1 = prob.controls_all_section(0)
        ur2 = prob.controls_all_section(1)
        ut1 = prob.controls_all_section(2)
        ut2 = prob.controls_all_section(3)
        return (ur1 + ur2) + (ut1 + ut2)

    t = prob.time_all_section
    G = api.Guess
    prob.set_states_all_section(0, G.linear(t, sc.r0, sc.rf))
    prob.set_states_all_section(1, G.linear(t, sc.vr0, sc.vrf))
    prob.set_states_all_section(2, G.linear(t, sc.vt0, sc.vtf))
    prob.set_controls_all_section(0, G.linear(t, sc.u_max, sc.u_max))
    prob.set_controls_all_section(2, G.linear(t, sc.u_max, sc.u_max))
    prob.dynamics = [dyn]
    prob.knot_states_smooth = []
    prob.cost = cost
    prob.running_cost = running
    prob.equality = eq
    prob.inequality = ineq

    def sanitize(prob, P):
        n0 = prob.nodes[0]
        P[:, 0:n0] = np.maximum(P[:, 0:n0], 0.5)            # r >= 0.5
        return P

    return Workload("low_thrust", prob, sc, sanitize)


# name -> (builder, default nodes) for the BASELINE.json configs
CONFIGS = {
    "cfg1_brachistochrone20": (brachistochrone, (20,)),
    "cfg2_goddard50": (goddard, (50,)),
    "cfg3_goddard_knot30x2": (goddard_knot, (30, 30)),
    "cfg4_polar3x40": (polar_3phase, (40, 40, 40)),
    "cfg5_lowthrust128": (low_thrust, (128,)),
    # the shipped example sizes (used by the parity tests against the examples)
    "ex05_goddard_knot25x2": (goddard_knot, (25, 25)),
    "ex09_polar_tsto20x2": (polar_tsto, (20, 20)),
    "ex10_lowthrust100": (low_thrust, (100,)),
}


def build(name, api):
    fn, nodes = CONFIGS[name]
    return fn(api, nodes)


def bounds_arrays(prob):
    """``prob.bounds`` (list of (lb|None, ub|None)) -> two float64 arrays with +-inf."""
    lb = np.array([-np.inf if b[0] is None else float(b[0]) for b in prob.bounds])
    ub = np.array([np.inf if b[1] is None else float(b[1]) for b in prob.bounds])
    return lb, ub


def make_batch(wl, B, first=0, seed0=SEED0):
    """Seeded synthetic instance batch: instance ``b`` is the
    shipped guess with 1 % multiplicative Gaussian jitter on every state/control
    entry and +-5 % uniform jitter on the final times, clipped into the bounds.
    One ``default_rng(seed0 + b)`` per instance so any sub-range is reproducible."""
    prob = wl.prob
    n = int(prob.number_of_variables)
    nsec = int(prob.number_of_section)
    p0 = np.asarray(prob.p, dtype=np.float64)
    lb, ub = bounds_arrays(prob)
    P = np.empty((B, n), dtype=np.float64)
    for i in range(B):
        rng = np.random.default_rng(seed0 + first + i)
        z = rng.standard_normal(n - nsec)
        u = rng.uniform(-1.0, 1.0, nsec)
        P[i, :n - nsec] = p0[:n - nsec] * (1.0 + 0.01 * z)
        P[i, n - nsec:] = p0[n - nsec:] * (1.0 + 0.05 * u)
    if wl.sanitize is not None:
        P = wl.sanitize(prob, P)
    np.clip(P, lb, ub, out=P)
    return P


# --------------------------------------------------------------------------
# edge-case problems for the parity tests (not benchmark configs)
# --------------------------------------------------------------------------
def two_stage_no_inequality(api, nodes=(9, 11)):
    """Two phases, linkage in the user equality, knot rows off, and an inequality function
    that returns NO rows (reference examples/07_Rocket_Ascent_TwoStage.py:81-103 does that)."""
    class Veh:
        GMe, Re, g0 = 3.986004418e14, 6371.0e3, 9.80665
        M0, M2, Mc, Cd, area, Isp = 5000, 2000, 0.4, 0.2, 10, 300.0

    veh = Veh()
    prob = api.Problem([0.0, 60.0, 150.0], list(nodes), [3, 3], [1, 1], 5)
    prob.set_unit_states_all_section(0, veh.Re)
    prob.set_unit_states_all_section(2, 1000.0)
    prob.set_unit_time(100.0)

    def dyn(prob, obj, section):
        R = prob.states(0, section)
        v = prob.states(1, section)
        m = prob.states(2, section)
        T = prob.controls(0, section)
        rho = 1.225 * np.exp(-(R - obj.Re) / 8500.0)
        drag = 0.5 * rho * v ** 2 * obj.Cd * obj.area
        d = api.Dynamics(prob, section)
        d[0] = v
        d[1] = (T - drag) / m - obj.GMe / R ** 2
        d[2] = - T / obj.g0 / obj.Isp
        return d()

    def eq(prob, obj):
        R = prob.states_all_section(0)
        v = prob.states_all_section(1)
        m = prob.states_all_section(2)
        r = api.Condition()
        r.equal(R[0], obj.Re, unit=prob.unit_states[0][0])
        r.equal(v[0], 0.0)
        r.equal(m[0], obj.M0)
        r.equal(m[-1], obj.M2 * obj.Mc)
        r.equal(prob.states(0, 0)[-1], prob.states(0, 1)[0], unit=prob.unit_states[0][0])
        r.equal(prob.states(1, 0)[-1], prob.states(1, 1)[0])
        r.equal(prob.states(2, 0)[-1], prob.states(2, 1)[0] + 1200)
        return r()

    def ineq(prob, obj):
        return api.Condition()()

    def cost(prob, obj):
        return -prob.states_all_section(0)[-1] / obj.Re

    t = prob.time_all_section
    prob.set_states_all_section(0, api.Guess.linear(t, veh.Re, veh.Re + 50e3))
    prob.set_states_all_section(1, api.Guess.linear(t, 0.0, 900.0))
    prob.set_states_all_section(2, api.Guess.linear(t, veh.M0, veh.M2 * veh.Mc))
    prob.set_controls_all_section(0, api.Guess.constant(t, 1.2 * veh.M0 * veh.g0))
    prob.set_controls_bounds_all_section(0, 0.0, 2.0 * veh.M0 * veh.g0)
    prob.dynamics = [dyn, dyn]
    prob.knot_states_smooth = [False]
    prob.cost = cost
    prob.equality = eq
    prob.inequality = ineq
    return Workload("two_stage_no_inequality", prob, veh, None)


def stress_mixed(api, nodes=(7, 33, 130)):
    """Synthetic stress case: three phases of very different sizes (130 nodes > the 128 the
    register-cached column code covers, so the generic code runs), a state-count change across
    knot 0 (no knot rows there) and smooth-state rows at knot 1, a running cost, units, bounds on
    states / controls / times, selects, several transcendental functions and a non-node-local row."""
    class Par:
        k1, k2, umax = 0.7, 1.3, 2.0

    par = Par()
    prob = api.Problem([0.0, 1.0, 2.5, 4.0], list(nodes), [2, 3, 3], [1, 2, 2], 3)
    prob.set_unit_states_all_section(0, 2.0)
    prob.set_unit_states(1, 1, 0.25)
    prob.set_unit_controls_all_section(0, 4.0)
    prob.set_unit_time(2.0)

    def dyn(prob, obj, section):
        x = prob.states(0, section)
        y = prob.states(1, section)
        u = prob.controls(0, section)
        d = api.Dynamics(prob, section)
        if section == 0:
            d[0] = y + obj.k1 * np.sin(x)
            d[1] = u - np.tanh(y) * x
        else:
            z = prob.states(2, section)
            w = prob.controls(1, section)
            sat = np.where(z > 0.5, 0.5 + 0.1 * (z - 0.5), z)
            d[0] = y * np.cos(z) - obj.k2 * x / (1.0 + x ** 2)
            d[1] = u - np.sqrt(1.0 + y ** 2) + np.abs(w)
            d[2] = w * np.exp(-sat) - np.arctan2(y, 1.0 + x ** 2)
        return d()

    def eq(prob, obj):
        r = api.Condition()
        r.equal(prob.states(0, 0)[0], 0.3, unit=prob.unit_states[0][0])
        r.equal(prob.states(1, 0)[0], -0.2)
        r.equal(prob.states(0, 1)[0], prob.states(0, 0)[-1], unit=prob.unit_states[0][0])
        r.equal(prob.states(1, 1)[0], prob.states(1, 0)[-1] * 1.5)
        r.equal(prob.states(2, 1)[0], 0.1)
        r.equal(prob.states(0, 2)[-1] ** 2 + prob.states(1, 2)[-1] ** 2, 1.0)
        r.equal(prob.time_final(1) - prob.time_final(0), 1.4, unit=prob.unit_time)
        return r()

    def ineq(prob, obj):
        u = prob.controls_all_section(0)
        r = api.Condition()
        r.lower_bound(u, -obj.umax, unit=prob.unit_controls[0][0])
        r.upper_bound(u, obj.umax, unit=prob.unit_controls[0][0])
        r.upper_bound(prob.controls(1, 2) ** 2 + prob.states(2, 2) ** 2, 9.0)
        r.lower_bound(prob.states(0, 1)[2:-1], -5.0)
        r.lower_bound(prob.time_final(0), 0.2)
        r.upper_bound(prob.states(1, 0) - prob.states(1, 0)[0], 6.0)      # not node-local -> expanded
        return r()

    def cost(prob, obj):
        return prob.time_final(-1) + 0.1 * prob.states(0, 2)[-1] ** 2

    def running(prob, obj):
        u = prob.controls_all_section(0)
        x = prob.states_all_section(0)
        return 0.5 * u ** 2 + 0.01 * np.cosh(0.1 * x)

    t = prob.time_all_section
    G = api.Guess
    prob.set_states_all_section(0, G.cubic(t, 0.3, 0.1, 0.8, 0.0))
    prob.set_states_all_section(1, G.linear(t, -0.2, 0.6))
    for s in (1, 2):
        prob.set_states(2, s, G.linear(prob.time[s], 0.1, 0.9))
        prob.set_controls(1, s, G.constant(prob.time[s], 0.3))
    prob.set_controls_all_section(0, G.linear(t, 0.5, -0.5))
    prob.set_states_bounds_all_section(0, -3.0, 3.0)
    prob.set_states_bounds(2, 2, 0.0, None)
    prob.set_controls_bounds_all_section(0, -par.umax, par.umax)
    prob.set_time_final_bounds(0, 0.2, 3.0)
    prob.set_time_final_bounds(2, None, 9.0)
    prob.dynamics = [dyn, dyn, dyn]
    prob.knot_states_smooth = [True, True]       # knot 0 is skipped anyway: state counts differ
    prob.cost = cost
    prob.running_cost = running
    prob.equality = eq
    prob.inequality = ineq
    return Workload("stress_mixed", prob, par, None)


CONFIGS["edge_two_stage_no_inequality"] = (two_stage_no_inequality, (9, 11))
CONFIGS["edge_stress_mixed"] = (stress_mixed, (7, 33, 130))
CONFIGS["edge_stress_small"] = (stress_mixed, (3, 4, 5))


def table_lookup(api, nodes=(12, 9)):
    """Edge case: data tables (`scipy.interpolate.interp1d`) inside the callbacks, the way
    reference examples/11_Polar_TSTO_Taiki.py:21-27,94-98 uses them: a float table with constant
    fill outside its range (SciPy takes the numpy.interp path), one with fill_value="extrapolate"
    (SciPy's two-term formula) and one with NaN fill that is only used in range."""
    from scipy import interpolate

    class Air:
        alt = np.array([0.0, 1.0, 2.5, 4.0, 7.0, 11.0, 15.0, 20.0, 32.0, 47.0])
        rho = np.array([1.225, 1.112, 0.957, 0.819, 0.590, 0.365, 0.194, 0.0880, 0.0132, 0.00143])
        snd = np.array([340.3, 336.4, 330.6, 324.6, 312.3, 295.1, 295.1, 295.1, 303.0, 329.8])
        mach = np.array([0.0, 0.4, 0.8, 1.0, 1.2, 2.0, 4.0])
        cd = np.array([0.30, 0.31, 0.38, 0.55, 0.62, 0.45, 0.33])
        density = interpolate.interp1d(alt, rho, bounds_error=False, fill_value=(rho[0], 0.0))
        sound = interpolate.interp1d(alt, snd, bounds_error=False, fill_value=(snd[0], snd[-1]))
        drag = interpolate.interp1d(mach, cd, fill_value="extrapolate")
        gain = interpolate.interp1d(np.array([-10.0, 0.0, 10.0, 60.0]), np.array([0.5, 1.0, 1.5, 1.2]))
        g0 = 9.80665e-3          # km / s^2

    air = Air()
    prob = api.Problem([0.0, 40.0, 120.0], list(nodes), [3, 3], [1, 1], 5)
    prob.set_unit_states_all_section(1, 0.5)
    prob.set_unit_time(50.0)

    def dyn(prob, obj, section):
        h = prob.states(0, section)          # km
        v = prob.states(1, section)          # km/s
        m = prob.states(2, section)
        T = prob.controls(0, section)
        mach = np.sqrt(v ** 2) * 1000.0 / obj.sound(h)
        q = 0.5 * obj.density(h) * (v * 1000.0) ** 2
        d = api.Dynamics(prob, section)
        d[0] = v
        d[1] = (T * obj.gain(h) - 1e-6 * q * obj.drag(mach)) / m - obj.g0
        d[2] = -T / (2.5 + 0.1 * section)
        return d()

    def eq(prob, obj):
        r = api.Condition()
        r.equal(prob.states(0, 0)[0], 0.0)
        r.equal(prob.states(1, 0)[0], 0.05)
        r.equal(prob.states(2, 0)[0], 1.0)
        r.equal(obj.density(prob.states(0, 1)[-1]), 0.02)        # a table inside a scalar row
        return r()

    def ineq(prob, obj):
        h = prob.states_all_section(0)
        v = prob.states_all_section(1)
        r = api.Condition()
        r.upper_bound(0.5 * obj.density(h) * (v * 1000.0) ** 2, 60000.0, unit=1000.0)
        r.lower_bound(prob.states_all_section(2), 0.2)
        return r()

    def cost(prob, obj):
        return -prob.states(2, 1)[-1]

    t = prob.time_all_section
    G = api.Guess
    prob.set_states_all_section(0, G.cubic(t, 0.0, 0.05, 55.0, 0.3))     # leaves the 0..47 km table range
    prob.set_states_all_section(1, G.linear(t, 0.05, 1.4))
    prob.set_states_all_section(2, G.linear(t, 1.0, 0.35))
    prob.set_controls_all_section(0, G.linear(t, 0.03, 0.005))
    prob.set_states_bounds_all_section(0, -5.0, None)
    prob.set_controls_bounds_all_section(0, 0.0, 0.05)
    prob.dynamics = [dyn, dyn]
    prob.knot_states_smooth = [True]
    prob.cost = cost
    prob.equality = eq
    prob.inequality = ineq
    return Workload("table_lookup", prob, air, None)


CONFIGS["edge_table_lookup"] = (table_lookup, (12, 9))


def all_ops(api, nodes=(9, 6)):
    """Edge case: every numpy ufunc / operator the tracer accepts that no other workload uses
    (tan, arcsin, arccos, arctan, sinh, log, log10, sign, floor, ceil, square, reciprocal, fabs,
    minimum, maximum, fmin, fmax, clip, hypot, exp2, log2, float_power, numpy.interp tables, non-integer
    and negative powers, scalar ** array, comparisons, logical_and / or / not, unary minus / plus,
    deg2rad) in dynamics, point rows, scalar rows, the
    cost and a running cost -- smooth arguments only, so the forward differences are well defined."""
    class Par:
        a, b = 0.8, 1.7
        tab_x = np.array([0.5, 1.0, 1.6, 2.2, 2.4, 3.5])
        tab_y = np.array([0.1, 0.7, 0.4, 0.9, 1.5, 1.1])

    par = Par()
    prob = api.Problem([0.0, 2.0, 3.0], list(nodes), [3, 3], [2, 2], 3)
    prob.set_unit_states_all_section(1, 3.0)
    prob.set_unit_controls_all_section(1, 0.5)
    prob.set_unit_time(1.5)

    def dyn(prob, obj, section):
        x = prob.states(0, section)          # in (0.2, 0.9)
        y = prob.states(1, section)          # in (1, 3)
        z = prob.states(2, section)          # in (-0.6, 0.6)
        u = prob.controls(0, section)
        w = prob.controls(1, section)
        d = api.Dynamics(prob, section)
        d[0] = (np.tan(0.5 * x) + np.arcsin(z) * np.arccos(0.5 * z) - np.sinh(z) + np.log(y) * np.log10(1.0 + y)
                + np.clip(u, -0.1, 0.45) * np.hypot(x, y) + np.exp2(z) * np.log2(y)
                + np.interp(y, obj.tab_x, obj.tab_y) - np.interp(z, obj.tab_x - 2.0, obj.tab_y, left=-1.0, right=4.0)
                + np.float_power(y, 1.25) + np.clip(w, None, 0.2))
        d[1] = (np.minimum(u, 0.3 * y) + np.maximum(w, -x) + np.fmin(x, 0.95) * np.fmax(z, -0.9)
                + np.square(x) * np.reciprocal(y) + np.fabs(z - 2.0) + y ** 0.5 + y ** -1.5 + 2.0 ** x
                + np.power(y, obj.a))
        d[2] = (np.arctan(obj.b * z) + np.sign(y) * np.floor(y + 10.25) * 0.01 + np.ceil(x - 7.5) * 0.02
                + np.where(np.logical_and(x > 0.0, y >= 0.5), -z, +z)
                + np.where(np.logical_or(x < -1.0, np.logical_not(y <= 100.0)), 5.0, np.deg2rad(u))
                + np.where(np.not_equal(x, 12.0), 1.0, 0.0) * np.where(np.equal(y, -3.0), 2.0, w))
        return d()

    def eq(prob, obj):
        r = api.Condition()
        r.equal(prob.states(0, 0)[0], 0.3)
        r.equal(np.log(prob.states(1, 0)[0]), np.log(1.5), unit=2.0)
        r.equal(np.tan(prob.states(2, 0)[0]), 0.0)
        r.equal(prob.states(0, 1)[0], prob.states(0, 0)[-1])
        r.equal(prob.states(1, 1)[0] ** 1.5, prob.states(1, 0)[-1] ** 1.5)
        r.equal(np.arctan(prob.states(2, 1)[0]), np.arctan(prob.states(2, 0)[-1]))
        return r()

    def ineq(prob, obj):
        r = api.Condition()
        r.lower_bound(np.log10(prob.states_all_section(1)), -1.0)
        r.upper_bound(np.sinh(prob.states_all_section(2)), 4.0)
        r.lower_bound(np.minimum(prob.controls_all_section(0), 1.0), -2.0)
        r.upper_bound(np.maximum(prob.time_final(0), 0.5), 9.0, unit=prob.unit_time)
        return r()

    def cost(prob, obj):
        return np.arcsin(0.5 * prob.states(2, 1)[-1]) + prob.time_final(-1) ** 1.25

    def running(prob, obj):
        return 0.1 * np.square(prob.controls_all_section(0)) + 0.01 * np.tan(0.3 * prob.states_all_section(0))

    t = prob.time_all_section
    G = api.Guess
    prob.set_states_all_section(0, G.linear(t, 0.3, 0.8))
    prob.set_states_all_section(1, G.cubic(t, 1.5, 0.2, 2.6, 0.0))
    prob.set_states_all_section(2, G.linear(t, -0.4, 0.5))
    prob.set_controls_all_section(0, G.linear(t, 0.2, 0.6))
    prob.set_controls_all_section(1, G.constant(t, 0.3))
    prob.set_states_bounds_all_section(0, 0.05, 0.95)
    prob.set_states_bounds_all_section(2, -0.9, 0.9)
    prob.dynamics = [dyn, dyn]
    prob.knot_states_smooth = [False]
    prob.cost = cost
    prob.running_cost = running
    prob.equality = eq
    prob.inequality = ineq
    return Workload("all_ops", prob, par, None)


CONFIGS["edge_all_ops"] = (all_ops, (9, 6))


def nonautonomous(api, nodes=(11, 8)):
    """Edge case: explicitly time-dependent callbacks, legal in the reference because it evaluates them eagerly
    (optimize.py:685): the dynamics read the phase's final / start time (`prob.time_final(s)`, :349-360), the
    LGL abscissae `prob.tau[s]` (per-node constants, :786-791) and a user table with one value per node; a path
    constraint and the running cost use `prob.time_update()` (:518-531).  On the device the final times become
    global inputs of the node programs, whose Jacobian columns are dense in the phases that read them."""
    class Obj:
        pass

    obj = Obj()
    rng = np.random.default_rng(12)
    obj.gust = [0.2 * rng.standard_normal(N) for N in nodes]
    prob = api.Problem([0.0, 1.5, 4.0], list(nodes), [2, 2], [1, 1], 5)
    # (a copy of the initial time grid: prob.time itself is overwritten by every time_update() call in the
    # reference, optimize.py:526-530, so reading it in one callback and calling time_update() in another would
    # make the result depend on the order SciPy calls them in)
    obj.ramp = [np.array(t, dtype=float) for t in prob.time]
    prob.set_unit_states_all_section(1, 2.0)
    prob.set_unit_time(2.0)

    def dyn(prob, obj, section):
        x = prob.states(0, section)
        v = prob.states(1, section)
        u = prob.controls(0, section)
        tf = prob.time_final(section)
        t0 = prob.time_start(section)
        t = (tf - t0) / 2.0 * prob.tau[section] + (tf + t0) / 2.0        # physical time at the nodes
        d = api.Dynamics(prob, section)
        d[0] = v + obj.gust[section] * np.sin(0.7 * t)
        d[1] = u - 0.3 * v * t / tf + 0.1 * x * obj.ramp[section]
        return d()

    def eq(prob, obj):
        r = api.Condition()
        r.equal(prob.states(0, 0)[0], 0.0)
        r.equal(prob.states(1, 0)[0], 0.5, unit=2.0)
        r.equal(prob.states(0, 1)[-1], 3.0)
        return r()

    def ineq(prob, obj):
        r = api.Condition()
        r.lower_bound(prob.states_all_section(0) + 0.05 * prob.time_update(), -5.0)
        r.upper_bound(prob.controls_all_section(0), 2.0)
        r.lower_bound(prob.time_final(0), 0.5)
        return r()

    def running(prob, obj):
        u = prob.controls_all_section(0)
        return u ** 2 * (1.0 + 0.1 * prob.time_update())

    def cost(prob, obj):
        return prob.time_final(-1)

    t = prob.time_all_section
    G = api.Guess
    prob.set_states_all_section(0, G.linear(t, 0.0, 3.0))
    prob.set_states_all_section(1, G.cubic(t, 0.5, 0.2, 1.0, -0.1))
    prob.set_controls_all_section(0, G.linear(t, 0.6, -0.2))
    prob.set_controls_bounds_all_section(0, -2.5, 2.5)
    prob.dynamics = [dyn, dyn]
    prob.knot_states_smooth = [True]
    prob.cost = cost
    prob.running_cost = running
    prob.equality = eq
    prob.inequality = ineq
    return Workload("nonautonomous", prob, obj, None)


CONFIGS["edge_nonautonomous"] = (nonautonomous, (11, 8))
CONFIGS["edge_nonautonomous_big"] = (nonautonomous, (40, 33))


def picked_dynamics(api, nodes=(9, 12)):
    """Edge case: the dynamics and the running cost read PICKED elements of the decision vector at every node --
    the phase's own initial state (`v[0]`, the way a mass ratio m / m[0] appears), its last control value, an
    element of the OTHER phase -- next to the final time.  Legal in the reference (eager evaluation,
    optimize.py:685); on the device these variables become global inputs of the node programs and their Jacobian
    columns are dense in the phases that read them."""
    class Obj:
        k = 0.35

    obj = Obj()
    prob = api.Problem([0.0, 1.0, 2.5], list(nodes), [2, 2], [1, 1], 5)
    prob.set_unit_states_all_section(0, 3.0)
    prob.set_unit_controls_all_section(0, 0.5)

    def dyn(prob, obj, section):
        x = prob.states(0, section)
        v = prob.states(1, section)
        u = prob.controls(0, section)
        v0 = v[0]                                     # picked: this phase's first node
        uN = u[-1]                                    # picked control: this phase's last node
        other = prob.states(0, 1 - section)[2]        # picked: an element of the other phase
        d = api.Dynamics(prob, section)
        d[0] = v * (1.0 + 0.1 * np.sin(v0)) + 0.05 * other
        d[1] = u - obj.k * v / (1.0 + v0 ** 2) + 0.02 * uN * x / prob.time_final(section)
        return d()

    def eq(prob, obj):
        r = api.Condition()
        r.equal(prob.states(0, 0)[0], 0.0)
        r.equal(prob.states(1, 0)[0], 0.4)
        r.equal(prob.states(0, 1)[-1], 2.0)
        return r()

    def ineq(prob, obj):
        r = api.Condition()
        r.upper_bound(prob.controls_all_section(0), 2.0)
        r.lower_bound(prob.states_all_section(1), -1.0)
        return r()

    def running(prob, obj):
        u = prob.controls_all_section(0)
        return u ** 2 * (1.0 + 0.3 * prob.states(1, 0)[0])

    def cost(prob, obj):
        return prob.time_final(-1)

    t = prob.time_all_section
    G = api.Guess
    prob.set_states_all_section(0, G.linear(t, 0.0, 2.0))
    prob.set_states_all_section(1, G.cubic(t, 0.4, 0.3, 0.9, -0.2))
    prob.set_controls_all_section(0, G.linear(t, 0.7, -0.1))
    prob.dynamics = [dyn, dyn]
    prob.knot_states_smooth = [True]
    prob.cost = cost
    prob.running_cost = running
    prob.equality = eq
    prob.inequality = ineq
    return Workload("picked_dynamics", prob, obj, None)


CONFIGS["edge_picked_dynamics"] = (picked_dynamics, (9, 12))
