"""Host-side mirror of the reference's solver API for the batched GPU path.

Names follow the reference (file:line under /root/reference):
  load_hsddp_setting      loadHSDDPSetting                 HSDDPSolver/common/HSDDP_CompoundTypes.h:57-82
  HKDProblem              HKDProblem<T>                    HKDMPC/HKD-TrajOpt/HKDProblem.h:93-168
  MHPCProblem             MHPCProblem<T>                   MHPC/MHPC-Trajopt/MHPCProblem.h:169-289
  MultiPhaseDDP           MultiPhaseDDP<T> (batched)       HSDDPSolver/header/MultiPhaseDDP.h:25-146
Everything numeric happens behind the C ABI (include/cafe_gpu.h); this file only marshals."""
import ctypes as C
import os

import numpy as np

from ._ctypes_defs import (CAFE_NKERNELS, CAFE_REF_W, CAFE_TRACE_W, MODEL_DIMS, Deck, Info, Options)
from .lib import check, lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(REPO, "data")


def load_hsddp_setting(fname):
    o = Options()
    check(lib.cafe_options_load(fname.encode(), C.byref(o)))
    return o


class _DeckOwner:
    def __init__(self):
        self._h = C.c_void_p()

    @property
    def deck(self):
        return lib.cafe_deck_get(self._h)

    def phases(self):
        d = self.deck.contents
        return [d.phase[i] for i in range(d.n_phases)]

    def reference_records(self):
        """Copy of the deck's per-knot reference records, [n_records, CAFE_REF_W] (layout: include/cafe_deck.h CAFE_REF_*)."""
        d = self.deck.contents
        return np.ctypeslib.as_array(d.ref, shape=(d.n_records, CAFE_REF_W)).copy()

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            lib.cafe_deck_free(self._h)
            self._h = C.c_void_p()


class HKDProblem(_DeckOwner):
    """set_problem_data(config) + initialization(): HKDProblem.cpp:15-111. Hard-coded plan values of
    HKDMPCSolver::initialize (HKDMPC.cpp:26-28) are the defaults."""

    def __init__(self, reference_csv, constraint_params=None, plan_duration=0.6, time_step=0.01,
                 nsteps_between_mpc=2, k0=0, mpc_update=False):
        """mpc_update: the deck stands for the problem after HKDProblem::update (HKDProblem.cpp:117-222) rather than after initialization():
        a tail phase of at most 2 knots was opened by that update and has no shooting states yet (:213-217)."""
        super().__init__()
        constraint_params = constraint_params or os.path.join(DATA, "HKDMPC/settings/constraint_params.info")
        check(lib.cafe_deck_build_hkd(reference_csv.encode(), constraint_params.encode(), plan_duration, time_step,
                                      nsteps_between_mpc, k0, C.byref(self._h)))
        self.single_shooting_phase = -1
        if mpc_update:
            r = C.c_int(-1)
            check(lib.cafe_deck_mark_mpc_update(self._h, 2, C.byref(r)))   # the literal 2 of HKDProblem.cpp:213
            self.single_shooting_phase = r.value

    def initial_state(self, body, qJ):
        """compute_hkd_state (HKDModel.h:66-96) with the first phase's contact."""
        body = np.ascontiguousarray(body, dtype=np.float64)
        qJ = np.ascontiguousarray(qJ, dtype=np.float64)
        x0 = np.zeros(24)
        dp = C.POINTER(C.c_double)
        check(lib.cafe_hkd_state(body.ctypes.data_as(dp), qJ.ctypes.data_as(dp), self.phases()[0].contact,
                                 x0.ctypes.data_as(dp)))
        return x0


class MHPCProblem(_DeckOwner):
    """MHPCProblem<T>::initialization (MHPCProblem.cpp:13-250) from mhpc_config.info."""

    def __init__(self, reference_csv, mhpc_config=None, settings_root=None, k0=0, mpc_update_nsteps=0):
        """mpc_update_nsteps > 0: the deck stands for the problem after MHPCProblem::update (shift of nsteps = dt_mpc / dt_wb knots)
        rather than after initialization(): a freshly opened tail phase (horizon <= nsteps) has no shooting states (MHPCProblem.cpp:366-369)."""
        super().__init__()
        mhpc_config = mhpc_config or os.path.join(DATA, "MHPC/settings/mhpc_config.info")
        settings_root = settings_root or DATA  # plays the role of the reference's "../"
        check(lib.cafe_deck_build_mhpc(reference_csv.encode(), mhpc_config.encode(), settings_root.encode(), k0,
                                       C.byref(self._h)))
        self.single_shooting_phase = -1
        if mpc_update_nsteps > 0:
            r = C.c_int(-1)
            check(lib.cafe_deck_mark_mpc_update(self._h, mpc_update_nsteps, C.byref(r)))
            self.single_shooting_phase = r.value


class LocoProblem(_DeckOwner):
    """LocoProblem<T> (Locomotion/LocoProblem.cpp:7-84; driver Loco_TO.cpp:16-82): whole-body-only locomotion trajectory
    optimisation from Locomotion/settings/loco_config.info — torque-limit and GRF barriers only, touchdown constraints as in MHPC."""

    def __init__(self, reference_csv=None, loco_config=None, settings_root=None, k0=0):
        super().__init__()
        reference_csv = reference_csv or os.path.join(DATA, "Reference/Data/flypace/quad_reference.csv")  # loco_config.info: referenceFile flypace
        loco_config = loco_config or os.path.join(DATA, "MHPC/MHPC-Trajopt/Locomotion/settings/loco_config.info")
        settings_root = settings_root or DATA
        check(lib.cafe_deck_build_loco(reference_csv.encode(), loco_config.encode(), settings_root.encode(), k0,
                                       C.byref(self._h)))


class BarrelRollProblem(_DeckOwner):
    """The in-place barrel roll of BarrelRollTO.cpp:65-275: six hand-scheduled whole-body phases, per-phase weights, fixed desired
    states, BarrelRoll:: barriers (incl. the joint-speed limit), four-foot touchdown constraints after both flight phases."""

    def __init__(self, cost_weights=None, constraint_params=None):
        super().__init__()
        d = os.path.join(DATA, "MHPC/MHPC-Trajopt/BarrelRoll/setting")
        cost_weights = cost_weights or os.path.join(d, "br_cost_weights.JSON")
        constraint_params = constraint_params or os.path.join(d, "br_constraint_params.info")
        check(lib.cafe_deck_build_barrel_to(cost_weights.encode(), constraint_params.encode(), C.byref(self._h)))

    def initial_guess(self, x0):
        """Packed guesses [B, solution_size] holding the interpolated state trajectory of BarrelRollTO.cpp:131-147 (for set_initial_guess)."""
        x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
        g = np.zeros((x0.shape[0], lib.cafe_solution_size(self.deck)))
        check(lib.cafe_barrel_to_initial_guess(self.deck, x0.ctypes.data_as(C.c_void_p), x0.shape[0], g.ctypes.data_as(C.c_void_p)))
        return g


LCM_FIELDS = (("torque", 12), ("eul", 3), ("pos", 3), ("qJ", 12), ("vWorld", 3), ("eulrate", 3), ("qJd", 12), ("GRF", 12), ("feedback", 432),
              ("Qu", 12), ("Quu", 144), ("Qux", 432))   # lcmtypes/MHPC_Command_lcmt.lcm, per-problem fields in struct order


def unpack_lcm_command(rec, n_steps):
    """Splits one float32 record of get_lcm_commands into {field: [n_steps, width]} (matrices stay column-major flat, as on the wire)."""
    out, off = {}, 0
    for name, w in LCM_FIELDS:
        out[name] = rec[off:off + n_steps * w].reshape(n_steps, w)
        off += n_steps * w
    return out


class MultiPhaseDDP:
    """Batched MultiPhaseDDP: set_multiPhaseProblem (constructor), set_initial_condition, solve,
    get_solver_info; results are packed host arrays instead of in-place Trajectory deques."""

    def __init__(self, problem, device=0, max_batch=1):
        self.problem = problem
        self._h = C.c_void_p()
        check(lib.cafe_gpu_create(problem.deck, device, max_batch, C.byref(self._h)))
        self.B = 0
        self.x0 = None

    def close(self):
        if self._h and self._h.value:
            lib.cafe_gpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_initial_condition(self, x0):
        x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
        self.x0 = x0
        self.B = x0.shape[0]

    def set_profiling(self, on):
        check(lib.cafe_gpu_set_profiling(self._h, 1 if on else 0))

    def solve(self, option):
        check(lib.cafe_gpu_solve_batch(self._h, self.x0.ctypes.data_as(C.c_void_p), self.B, C.byref(option)))

    def solve_device(self, x0_dev_ptr, ldx, B, option):
        self.B = B
        check(lib.cafe_gpu_solve_batch_device(self._h, C.c_void_p(x0_dev_ptr), ldx, B, C.byref(option)))

    def get_solver_info(self):
        info = (Info * self.B)()
        check(lib.cafe_gpu_get_info(self._h, info))
        return [i.as_dict() for i in info]

    def get_history(self, cap=64):
        h = np.zeros((self.B, cap, 4))
        check(lib.cafe_gpu_get_history(self._h, h.ctypes.data_as(C.c_void_p), cap))
        return h

    def get_trace(self, cap=64):
        t = np.zeros((self.B, cap, CAFE_TRACE_W))
        check(lib.cafe_gpu_get_trace(self._h, t.ctypes.data_as(C.c_void_p), cap))
        return t

    def solution_size(self):
        return lib.cafe_solution_size(self.problem.deck)

    def get_solution(self, b0=0, nb=None):
        nb = self.B - b0 if nb is None else nb
        s = np.zeros((nb, self.solution_size()))
        check(lib.cafe_gpu_get_solution(self._h, b0, nb, s.ctypes.data_as(C.c_void_p)))
        return s

    def get_commands(self, n_gain_knots=8, out=None):
        sz = lib.cafe_command_size(self.problem.deck, n_gain_knots)
        if out is None:
            out = np.zeros((self.B, sz))
        check(lib.cafe_gpu_get_commands(self._h, n_gain_knots, out.ctypes.data_as(C.c_void_p)))
        return out

    def set_references(self, refs):
        """Per-problem reference records [B, n_records, CAFE_REF_W] on the shared phase schedule (None: back to the deck's)."""
        if refs is None:
            check(lib.cafe_gpu_set_references(self._h, None, 0))
            return
        refs = np.ascontiguousarray(refs, dtype=np.float64)
        check(lib.cafe_gpu_set_references(self._h, refs.ctypes.data_as(C.c_void_p), refs.shape[0]))

    def set_initial_guess(self, guess):
        """Warm start: packed solutions [B, solution_size] whose Xbar / Ubar / K start the next solves (None: cold start)."""
        if guess is None:
            check(lib.cafe_gpu_set_initial_guess(self._h, None, 0))
            return
        guess = np.ascontiguousarray(guess, dtype=np.float64)
        assert guess.shape[1] == lib.cafe_solution_size(self.problem.deck)
        check(lib.cafe_gpu_set_initial_guess(self._h, guess.ctypes.data_as(C.c_void_p), guess.shape[0]))

    def set_al_params(self, al):
        """Augmented-Lagrangian parameters the next solves start from, [B, n_phases, 4, 2] = (sigma, lambda) per touchdown-constraint element
        (None: the deck's TD_AL values). The reference carries them from one MPC step to the next (cafe_gpu.h); update_deck and
        shift_guess_from do so on the device by themselves."""
        if al is None:
            check(lib.cafe_gpu_set_al_params(self._h, None, 0))
            return
        al = np.ascontiguousarray(al, dtype=np.float64)
        assert al.shape[1:] == (len(self.problem.phases()), 4, 2)
        check(lib.cafe_gpu_set_al_params(self._h, al.ctypes.data_as(C.c_void_p), al.shape[0]))

    def get_al_params(self, B=None):
        """(sigma, lambda) of every touchdown-constraint element as the last solve left them: [B, n_phases, 4, 2]."""
        B = B or self.B
        out = np.zeros((B, len(self.problem.phases()), 4, 2))
        check(lib.cafe_gpu_get_al_params(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def shift_guess_from(self, prev, prev_k0, k0, B=None):
        """Warm start from the solution held by the solver `prev` (deck at start offset prev_k0), shifted to this solver's deck (start
        offset k0) on the device - the receding-horizon update of MHPCProblem::update without a host round trip."""
        check(lib.cafe_gpu_shift_guess(self._h, prev._h, prev_k0, k0, self.B if B is None else B))

    def update_deck(self, problem, k_advance, B=None):
        """The MPC update on this solver (cafe_gpu_update_deck): `problem` (the deck re-cut k_advance knots later) replaces the current one,
        the previous solution becomes the warm start, device buffers are re-used. B = 0: cold start on the new deck."""
        check(lib.cafe_gpu_update_deck(self._h, problem.deck, k_advance, self.B if B is None else B))
        self.problem = problem

    def planned_state(self, knots_ahead):
        """[B, n] planned state `knots_ahead` knots after the start of the plan."""
        n = MODEL_DIMS[self.problem.phases()[0].model][0]
        out = np.zeros((self.B, n))
        check(lib.cafe_gpu_get_planned_state(self._h, knots_ahead, out.ctypes.data_as(C.c_void_p)))
        return out

    def get_lcm_commands(self, n_steps=8, out=None):
        """float32 MHPC_Command_lcmt record per problem (see include/cafe_gpu.h); use unpack_lcm_command to name the fields.
        out: caller's [B, cafe_lcm_command_size] float32 buffer (page-locked memory makes the copy several times faster)."""
        if out is None:
            out = np.zeros((self.B, lib.cafe_lcm_command_size(n_steps)), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (self.B, lib.cafe_lcm_command_size(n_steps))
        check(lib.cafe_gpu_get_lcm_commands(self._h, n_steps, out.ctypes.data_as(C.c_void_p)))
        return out

    def get_hkd_lcm_commands(self, n_steps=9):
        """Per-problem float32 fields of hkd_command_lcmt (HKDMPC.cpp:243-290): [B, 180 n_steps] = hkd_controls[N][24], des_body_state[N][12],
        feedback[N][12][12]; n_steps = nsteps_between_mpc + 7 in the reference."""
        out = np.zeros((self.B, lib.cafe_hkd_lcm_command_size(n_steps)), dtype=np.float32)
        check(lib.cafe_gpu_get_hkd_lcm_commands(self._h, n_steps, out.ctypes.data_as(C.c_void_p)))
        return out

    def get_lcm_commands_device(self, n_steps, dev_ptr):
        check(lib.cafe_gpu_get_lcm_commands_device(self._h, n_steps, C.c_void_p(dev_ptr)))

    def debug_get(self, name, phase, b=0):
        buf = np.zeros(64 * 36 * 36 + 64)
        n = lib.cafe_gpu_debug_get(self._h, name.encode(), phase, b, buf.ctypes.data_as(C.c_void_p))
        if n < 0:
            check(int(n))
        return buf[:n].copy()

    def command_size(self, n_gain_knots=8):
        return lib.cafe_command_size(self.problem.deck, n_gain_knots)

    def get_commands_async(self, n_gain_knots, out_pinned, slot=0):
        """cafe_gpu_get_commands_async: out_pinned = page-locked [B, command_size] array; returns at once, commands_wait(slot) blocks until it holds the records."""
        check(lib.cafe_gpu_get_commands_async(self._h, n_gain_knots, out_pinned.ctypes.data_as(C.c_void_p), slot))

    def commands_wait(self, slot=0):
        check(lib.cafe_gpu_commands_wait(self._h, slot))

    def gather_commands_async(self, n_gain_knots, per_rank, out_dev_ptr, out_host_ptr, slot=0):
        check(lib.cafe_gpu_gather_commands_async(self._h, n_gain_knots, per_rank, C.c_void_p(out_dev_ptr), C.c_void_p(out_host_ptr), slot))

    def get_commands_device(self, n_gain_knots, dev_ptr):
        check(lib.cafe_gpu_get_commands_device(self._h, n_gain_knots, C.c_void_p(dev_ptr)))

    # ---- one process per GPU (torchrun / MPI): NCCL communicator of this rank's solver and the final gather (include/cafe_gpu.h)
    def comm_init_rank(self, nranks, rank, unique_id):
        check(lib.cafe_gpu_comm_init_rank(self._h, nranks, rank, unique_id))

    def gather_commands(self, n_gain_knots, per_rank, out_dev_ptr):
        check(lib.cafe_gpu_gather_commands(self._h, n_gain_knots, per_rank, C.c_void_p(out_dev_ptr)))

    def solve_ms(self):
        v = C.c_double()
        check(lib.cafe_gpu_get_solve_ms(self._h, C.byref(v)))
        return v.value

    def get_timing(self):
        ms = (C.c_double * CAFE_NKERNELS)()
        n = (C.c_long * CAFE_NKERNELS)()
        ticks = C.c_int()
        check(lib.cafe_gpu_get_timing(self._h, C.byref(ms), C.byref(n), C.byref(ticks)))
        names = ["roll", "select", "accept", "lq", "bwd", "misc", "wb_terms", "wb_fwd", "wb_derivs", "wb_sens", "wb_cost"]
        un = (C.c_double * CAFE_NKERNELS)()
        check(lib.cafe_gpu_get_units(self._h, C.byref(un)))
        return {"ms": dict(zip(names, list(ms))), "launches": dict(zip(names, list(n))), "units": dict(zip(names, list(un))), "ticks": ticks.value}


def nccl_unique_id():
    """128-byte NCCL id made by rank 0; the launcher (torch.distributed, MPI, a file) hands it to every rank."""
    buf = C.create_string_buffer(128)
    check(lib.cafe_gpu_nccl_unique_id(buf))
    return buf.raw


def shard_range(B, nranks, rank):
    lo, hi = C.c_int(), C.c_int()
    check(lib.cafe_gpu_shard_range(B, nranks, rank, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


class MultiGPUDDP:
    """One process, several GPUs of one box (cafe_gpu_create_multi): the batch is cut into contiguous slices, one solver per GPU, the
    command records are gathered on the first GPU by one NCCL group of sends / receives."""

    def __init__(self, problem, ndev, max_batch, devices=None):
        self.problem = problem
        self._m = C.c_void_p()
        devs = (C.c_int * ndev)(*devices) if devices is not None else None
        check(lib.cafe_gpu_create_multi(problem.deck, ndev, devs, max_batch, C.byref(self._m)))
        self.ndev = ndev
        self.B = 0

    def close(self):
        if self._m and self._m.value:
            lib.cafe_gpu_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, x0, option):
        x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
        self.B = x0.shape[0]
        check(lib.cafe_gpu_multi_solve_batch(self._m, x0.ctypes.data_as(C.c_void_p), self.B, C.byref(option)))

    def update_deck(self, problem, k_advance, B=None):
        """cafe_gpu_multi_update_deck: the MPC update on every GPU's solver (B = 0: cold start on the new deck)"""
        check(lib.cafe_gpu_multi_update_deck(self._m, problem.deck, k_advance, self.B if B is None else B))
        self.problem = problem

    def get_solver_info(self):
        info = (Info * self.B)()
        check(lib.cafe_gpu_multi_get_info(self._m, info))
        return [i.as_dict() for i in info]

    def get_commands(self, n_gain_knots=8, out=None):
        sz = lib.cafe_command_size(self.problem.deck, n_gain_knots)
        if out is None:
            out = np.zeros((self.B, sz))
        check(lib.cafe_gpu_multi_get_commands(self._m, n_gain_knots, out.ctypes.data_as(C.c_void_p)))
        return out


def unpack_solution(deck, sol):
    """Split one packed solution (cafe_solution_size doubles) into named per-phase arrays."""
    out = []
    off = 0
    d = deck.contents
    for i in range(d.n_phases):
        ph = d.phase[i]
        n, m, p = MODEL_DIMS[ph.model]
        h = ph.horizon
        r = {}
        for name, shape in (("Xbar", (h + 1, n)), ("Ubar", (h, m)), ("Y", (h, p)), ("dU", (h, m)), ("K", (h, n, m)),
                            ("Qu", (h, m)), ("Quu", (h, m, m)), ("Qux", (h, n, m)), ("G", (h + 1, n))):
            cnt = int(np.prod(shape))
            a = np.asarray(sol[off:off + cnt]).reshape(shape)
            if len(shape) == 3:
                a = a.transpose(0, 2, 1)  # stored column-major per knot
            r[name] = a
            off += cnt
        out.append(r)
    return out


def measure_fp64_peak(device=0):
    v = C.c_double()
    check(lib.cafe_gpu_measure_fp64_peak(device, C.byref(v)))
    return v.value
