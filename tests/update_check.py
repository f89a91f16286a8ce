"""TEST INFRASTRUCTURE: an independent restatement of what the reference's MPC update does to the MHPC problem data, step by step, with the
reference's deque operations and float time arithmetic:
    MHPCProblem::update / update_WB_plan / update_SRB_plan                  MHPC/MHPC-Trajopt/MHPCProblem.cpp:252-397
    QuadReference::step                                                     Reference/QuadReference.cpp:33-52
    Trajectory::pop_front / push_back_state, SinglePhase::pop_front / push_back_default
                                                                            HSDDPSolver/source/TrajectoryManagement.cpp:130-228, SinglePhase.cpp:513-528
    is_phase_reach_end is false for every phase after initialisation (MHPCProblem.cpp:127)
It keeps the schedule (start / end times, horizons, contacts, reach-end flags, which phases carry a touchdown constraint, which have a
shooting-state set) and, per whole-body phase, the lists X / U / K the next solve starts from. Shares no code with csrc/host,
cafe_mpc_b200/mpc.py or include/hsddp_facade (the three places of the product that implement the update)."""
import numpy as np

from deck_check import Reference, approx_eq, approx_leq, expected_mhpc_deck, f32, read_info


class MHPCPlan:
    def __init__(self, csv, config_info, settings_root, k0=0):
        cfg = read_info(config_info)
        self.csv, self.k0 = csv, k0
        self.plan_wb, self.dt_wb = float(cfg["config.plan_dur_wb"]), float(cfg["config.dt_wb"])
        self.plan_srb, self.dt_srb = float(cfg["config.plan_dur_srb"]), float(cfg["config.dt_srb"])
        self.dt_mpc = f32(cfg["config.dt_mpc"])
        exp = expected_mhpc_deck(csv, config_info, settings_root, k0)
        wb = [p for p in exp if p["model"] == 1]
        self.start = [p["start"] for p in wb]; self.end = [p["end"] for p in wb]
        self.horizon = [p["horizon"] for p in wb]; self.contact = [p["contact"] for p in wb]
        self.reach_end = [False] * len(wb)                                   # quirk 14
        self.tconstr = [len(p["td_foot"]) > 0 for p in wb]                   # add_tconstr_one_phase at initialisation (look-ahead contact)
        self.has_ss = [True] * len(wb)                                       # update_SS_config(h + 1) at initialisation (:209)
        self.t_cur, self.k_cur = f32(0), 0                                   # QuadReference after initialize()
        self.ref_dt = f32(0.01)
        self.ref = Reference(csv, k0, f32(self.plan_wb + self.plan_srb))     # the sliding window: rows k_cur .. k_cur + sz
        self._all = Reference(csv, k0, 1e4)                                  # the whole file behind it
        # trajectories the first solve starts from: X = Xbar = reference states, U = K = 0 (:186-192)
        self.X = [[r["xr"].copy() for r in p["records"]] for p in wb]
        self.U = [[np.zeros(12) for _ in range(p["horizon"])] for p in wb]
        self.K = [[np.zeros((12, 36)) for _ in range(p["horizon"] + 1)] for p in wb]

    def load_solution(self, phases):
        """after a solve: X (= Xbar after update_nominal_trajectory), Ubar, K of the whole-body phases (K[h] stays zero)"""
        for i in range(len(self.horizon)):
            h = self.horizon[i]
            self.X[i] = [np.array(phases[i]["Xbar"][k]) for k in range(h + 1)]
            self.U[i] = [np.array(phases[i]["Ubar"][k]) for k in range(h)]
            self.K[i] = [np.array(phases[i]["K"][k]) for k in range(h)] + [np.zeros((12, 36))]

    def contact_at(self, t_rel):
        return tuple(int(c) for c in self._all.rows[self.k_cur + self.ref.index(t_rel)]["contact"])

    def update(self):
        # ---- quad_reference->step(dt_mpc)
        i = 1
        while approx_leq(f32(i) * self.ref_dt, self.dt_mpc):
            self.k_cur += 1; self.t_cur = f32(self.t_cur + self.ref_dt); i += 1
        nsteps = int(round(float(self.dt_mpc) / self.dt_wb))
        new_start_time = self.t_cur
        # ---- front end
        for _ in range(nsteps):
            first = f32(float(self.start[0]) + self.dt_wb)
            if approx_eq(self.end[0], first):
                for l in (self.start, self.end, self.horizon, self.contact, self.reach_end, self.tconstr, self.has_ss, self.X, self.U, self.K):
                    l.pop(0)
            else:
                self.X[0].pop(0); self.U[0].pop(0); self.K[0].pop(0)
                self.horizon[0] -= 1; self.start[0] = first
        # ---- back end
        for _ in range(nsteps):
            new_end = f32(float(self.end[-1]) + self.dt_wb)
            new_contact = self.contact_at(f32(new_end - new_start_time))
            change = new_contact != self.contact[-1]
            if change and self.reach_end[-1]:
                self.start.append(self.end[-1]); self.end.append(new_end); self.horizon.append(1); self.reach_end.append(False)
                self.contact.append(new_contact); self.tconstr.append(False); self.has_ss.append(False)
                self.X.append([np.zeros(36), np.zeros(36)]); self.U.append([np.zeros(12)]); self.K.append([np.zeros((12, 36)), np.zeros((12, 36))])
            else:
                self.end[-1] = new_end; self.horizon[-1] += 1
                if change:
                    self.reach_end[-1] = True; self.tconstr[-1] = True
                self.X[-1].append(self.X[-1][-1].copy()); self.U[-1].append(np.zeros(12)); self.K[-1].append(np.zeros((12, 36)))
        n = len(self.horizon)
        for i in range(n):
            if i < n - 1 or self.horizon[i] > nsteps:
                self.has_ss[i] = True
        return nsteps

    def touchdown_feet(self, i):
        """feet of the touchdown constraint phase i carries (add_tconstr_one_phase, :566-601), () if it carries none"""
        if not self.tconstr[i]:
            return ()
        nxt = self.contact[i + 1] if i < len(self.horizon) - 1 else self.contact_at(f32(self.plan_wb + float(self.dt_mpc)))
        return tuple(l for l in range(4) if self.contact[i][l] == 0 and nxt[l] == 1)


class HKDPlan:
    """HKDProblem::update (HKDMPC/HKD-TrajOpt/HKDProblem.cpp:117-222): one reference step of dt_sim per inner iteration, front pop / phase removal,
    tail growth / phase opening with the same reach-end rule (false for every phase after initialization(), :57-58: contact_prev is compared with
    itself), touchdown constraint attached when the change reaches the tail (:188-196), shooting states for every phase but a tail phase of at most
    two knots (:213-217), Ubar[0] of the front trajectory zeroed (:220). The schedule after initialization() is taken as an argument (it is pinned
    by the schedule goldens); everything after it is restated here."""

    def __init__(self, csv, k0, plan_duration, dt_sim, nsteps_between_mpc, horizons, contacts, td_feet):
        self.dt_sim, self.nsteps, self.plan = f32(dt_sim), nsteps_between_mpc, f32(plan_duration)
        self.dt_mpc = f32(f32(dt_sim) * f32(nsteps_between_mpc))
        self.horizon, self.contact = list(horizons), [tuple(c) for c in contacts]
        t, self.start, self.end = f32(0), [], []
        for h in horizons:
            self.start.append(t); t = f32(t + f32(h) * self.dt_sim); self.end.append(t)
        self.reach_end = [False] * len(horizons)
        self.td = [tuple(f) for f in td_feet]                    # frozen when add_tconstr_one_phase ran
        self.has_ss = [True] * len(horizons)
        self.t_cur, self.k_cur, self.ref_dt = f32(0), 0, f32(0.01)
        self.ref = Reference(csv, k0, self.plan)
        self._all = Reference(csv, k0, 1e4)
        self.X = self.U = self.K = None

    def load_solution(self, phases):
        n = len(self.horizon)
        self.X = [[np.array(phases[i]["Xbar"][k]) for k in range(self.horizon[i] + 1)] for i in range(n)]
        self.U = [[np.array(phases[i]["Ubar"][k]) for k in range(self.horizon[i])] for i in range(n)]
        self.K = [[np.array(phases[i]["K"][k]) for k in range(self.horizon[i])] + [np.zeros((24, 24))] for i in range(n)]

    def contact_at(self, t_rel):
        c = self._all.rows[self.k_cur + self.ref.index(t_rel)]["contact"]
        return (int(c[1]), int(c[0]), int(c[3]), int(c[2]))     # the HKD application flips right and left legs (QuadReference.cpp:372-406)

    def update(self):
        for _ in range(self.nsteps):
            i = 1
            while approx_leq(f32(i) * self.ref_dt, self.dt_sim):          # quad_ref_ptr->step(dt_sim)
                self.k_cur += 1; self.t_cur = f32(self.t_cur + self.ref_dt); i += 1
            new_start, new_end = self.t_cur, f32(self.t_cur + self.plan)
            self.start[0] = f32(self.start[0] + self.dt_sim)
            if approx_leq(self.end[0], new_start):
                for l in (self.start, self.end, self.horizon, self.contact, self.reach_end, self.td, self.has_ss, self.X, self.U, self.K):
                    l.pop(0)
            else:
                self.X[0].pop(0); self.U[0].pop(0); self.K[0].pop(0)
                self.horizon[0] -= 1; self.start[0] = new_start
            new_contact = self.contact_at(f32(new_end - new_start))
            change = new_contact != self.contact[-1]
            if change and self.reach_end[-1]:
                h = int(round(float(f32(new_end - self.end[-1]) / self.dt_sim)))
                self.start.append(self.end[-1]); self.end.append(new_end); self.horizon.append(h); self.reach_end.append(False)
                self.contact.append(new_contact); self.td.append(()); self.has_ss.append(False)
                self.X.append([np.zeros(24) for _ in range(h + 1)]); self.U.append([np.zeros(24) for _ in range(h)])
                self.K.append([np.zeros((24, 24)) for _ in range(h + 1)])
            else:
                self.end[-1] = new_end; self.horizon[-1] += 1
                if change:
                    self.reach_end[-1] = True
                self.X[-1].append(self.X[-1][-1].copy()); self.U[-1].append(np.zeros(24)); self.K[-1].append(np.zeros((24, 24)))
            if self.reach_end[-1]:
                nxt = self.contact_at(f32(self.plan + self.dt_mpc))
                self.td[-1] = tuple(l for l in range(4) if self.contact[-1][l] == 0 and nxt[l] == 1)
        n = len(self.horizon)
        for i in range(n):
            if i < n - 1 or self.horizon[i] > 2:
                self.has_ss[i] = True
        self.U[0][0] = np.zeros(24)
