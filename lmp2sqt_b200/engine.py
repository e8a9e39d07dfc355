"""Python mirror of the reference's call seam over the C-ABI (include/sqt.h).

The reference drives its hot path with four calls::

    InitSpacetimeCorr(cl)      lmp2sqt.c:446     once per input file
    EvalSpacetimeCorr(cl)      lmp2sqt.c:1599    once per frame, cl->snap = ReadDump(fp)
    PrintProcess(cl)           lmp2sqt.c:338     when countCorrAv == limitCorrAv, and at end of run

:class:`SpacetimeCorr` exposes exactly those under the same names (plus pythonic aliases), with
the same argument meaning and the same error behaviour (a malformed frame stops the file; a
failed allocation aborts), so parity tests read like the reference's main loop.
All numerics run in the CUDA engine; this module only moves pointers.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import SQT_FLUSH, SqtDeck, SqtFlags, SqtParams, SqtResults

NAMES2D = ["avF_qq2", "StdDevF_qq2", "avF_s_qq2", "StdDevF_s_qq2", "avF_d_qq2", "StdDevF_d_qq2", "avDrTable",
           "avC2_v_rho", "avC2_rho_v", "avC2_v_v",           # -V only (zero otherwise)
           "avC2_v_mu", "avC2_mu_v", "avC2_mu_rho", "avC2_rho_mu", "avC2_mu_mu"]   # -M (the first two: -V -M) only
NAMES1D = ["rrMSDAv", "rrMQDAv", "rrMSDCMAv", "rrCvvAv", "rrCvcmvcmAv", "rrCmmAv"]
NAMES_TENSOR = {"rrMSR1_R_Av": 3, "rrMSR2_VR_Av": 9}         # -s only: the reference's VecR3[T] / Rank2R3[T]


def _flag_ints(flags: str):
    """getopt flags of the reference (lmp2sqt.c:152-191) as (f, F, t, V, M, s): -F implies -f, -s implies -t."""
    f_F = "F" in flags
    return (int(f_F or "f" in flags), int(f_F), int("t" in flags or "s" in flags), int("V" in flags), int("M" in flags),
            int("s" in flags))


def _ncomp(flags: str) -> int:
    """arrays per SoA frame: x y z, then vx vy vz (-V or -s), then mux muy muz (-M)"""
    return 3 + (3 if ("V" in flags or "s" in flags) else 0) + (3 if "M" in flags else 0)


def _fill_stage(st: np.ndarray, flags: str, xyz, vel, mu) -> None:
    st[:3] = np.asarray(xyz, dtype=np.float64).T
    c = 3
    if "V" in flags or "s" in flags:
        if vel is None:
            raise SqtError("-V / -s need velocities")
        st[c:c + 3] = np.asarray(vel, dtype=np.float64).T
        c += 3
    if "M" in flags:
        if mu is None:
            raise SqtError("-M needs dipoles")
        st[c:c + 3] = np.asarray(mu, dtype=np.float64).T


class SqtError(RuntimeError):
    pass


@dataclass
class Deck:
    """in.lmp2sqt.* (lmp2sqt.c:82-92)."""
    kVal: float = 0.0981748
    deltaT: float = 0.06
    mass: float = 1.0
    rVal: float = 0.2
    limitCorrAv: int = 0
    nCBuffer: int = 150
    nCSpatial: int = 30
    nCTime: int = 150
    nCSkip: int = 0

    def to_c(self) -> SqtDeck:
        d = SqtDeck(self.kVal, self.deltaT, self.mass, self.rVal, self.limitCorrAv, self.nCBuffer, self.nCSpatial,
                    self.nCTime, self.nCSkip)
        for i in range(9):
            d.status[i] = 1
        return d

    @staticmethod
    def read(path: str) -> "Deck":
        d = SqtDeck()
        rc = _lib.host_lib().sqt_deck_read(os.fsencode(path), C.byref(d))
        if rc < 0:
            raise FileNotFoundError(path)
        return Deck(d.kVal, d.deltaT, d.mass, d.rVal, d.limitCorrAv, d.nCBuffer, d.nCSpatial, d.nCTime, d.nCSkip)


class Results:
    """Owner of an ``sqt_results`` (raw sums as the engine returns them; ``normalise`` = prePrintProcess)."""

    def __init__(self, nCSpatial: int, nCTime: int, n_types: int = 0):
        self.c = SqtResults()
        rc = _lib.engine_lib().sqt_results_alloc(C.byref(self.c), nCSpatial, nCTime)
        if rc:
            raise MemoryError("sqt_results_alloc")
        self.S, self.T = nCSpatial, nCTime
        self.n_types = n_types if n_types > 1 else 0
        if self.n_types and _lib.engine_lib().sqt_results_alloc_types(C.byref(self.c), n_types):
            raise MemoryError("sqt_results_alloc_types")

    def type_arrays(self) -> dict:
        """Per-type partials (raw sums): F_ab [a][b][S][T], F_s_a / DrTable_a [a][S][T], MSD_a / MQD_a [a][T], N_a [a]."""
        nt, S, T = self.n_types, self.S, self.T
        if not nt:
            return {}
        g = lambda name, shape: np.ctypeslib.as_array(getattr(self.c, name), shape=shape).copy()
        return {"avF_ab": g("avF_ab", (nt, nt, S, T)), "avF_s_a": g("avF_s_a", (nt, S, T)), "avDrTable_a": g("avDrTable_a", (nt, S, T)),
                "rrMSD_a": g("rrMSD_a", (nt, T)), "rrMQD_a": g("rrMQD_a", (nt, T)), "nPtls_a": np.array(list(self.c.nPtls_a)[:nt])}

    def __del__(self):
        try:
            _lib.engine_lib().sqt_results_free(C.byref(self.c))
        except Exception:
            pass

    def array(self, name: str) -> np.ndarray:
        p = getattr(self.c, name)
        shape = (self.S, self.T) if name in NAMES2D else ((self.T, NAMES_TENSOR[name]) if name in NAMES_TENSOR else (self.T,))
        return np.ctypeslib.as_array(p, shape=shape)

    def to_dict(self) -> dict:
        d = {n: self.array(n).copy() for n in NAMES2D + NAMES1D + list(NAMES_TENSOR)}
        for k in ("countCorrAv", "limitCorrAv", "nPtls", "g_Vol", "kVal", "first_Vol", "first_limitCorrAv", "number_call_print"):
            d[k] = getattr(self.c, k)
        d.update(self.type_arrays())
        return d

    def normalise(self, deck: Deck, flags: str = "Ft") -> None:
        """prePrintProcess (lmp2sqt.c:576-675), in place."""
        dk = deck.to_c()
        fl = SqtFlags(*_flag_ints(flags))
        _lib.host_lib().sqt_host_normalise_x(C.byref(self.c), C.byref(dk), C.byref(fl))

    def normalise_types(self, deck: Deck) -> None:
        dk = deck.to_c()
        _lib.host_lib().sqt_host_normalise_types(C.byref(self.c), C.byref(dk))

    def print_process(self, outdir: str, deck: Deck, flags: str = "Ft", date: str | None = None) -> None:
        """PrintProcess (lmp2sqt.c:338-367): normalise in place and write the reference's files into ``outdir``."""
        dk = deck.to_c()
        fl = SqtFlags(*_flag_ints(flags))
        rc = _lib.host_lib().sqt_host_print_process_x(os.fsencode(outdir), C.byref(self.c), C.byref(dk), C.byref(fl),
                                                      date.encode() if date else None)
        if rc:
            raise OSError("sqt_host_print_process failed in %s" % outdir)


class SpacetimeCorr:
    """One GPU's share of the space-time correlation state (the reference's MakeSqtClass + globals)."""

    def __init__(self, deck: Deck, flags: str = "Ft", device: int = 0, n_types: int = 0):
        self.lib = _lib.engine_lib()
        self.deck = deck
        self.flags = flags
        self.n_types = n_types if n_types > 1 else 0
        ff, fF, ft, fV, fM, fs = _flag_ints(flags)
        p = SqtParams(deck.kVal, deck.deltaT, deck.mass, deck.rVal, deck.limitCorrAv, deck.nCBuffer, deck.nCSpatial,
                      deck.nCTime, deck.nCSkip, ff, fF, ft, device, 0, fV, fM, fs, n_types)
        self.ncomp = _ncomp(flags)
        h = C.c_void_p()
        rc = self.lib.sqt_create(C.byref(p), C.byref(h))
        if rc:
            raise SqtError("sqt_create failed (%d): %s" % (rc, self.lib.sqt_last_error(None).decode()))
        self.h = h
        self.S, self.T = deck.nCSpatial, deck.nCTime
        self.B = max(1, min(deck.nCBuffer, deck.nCTime))
        self.device = device

    # ---- lifecycle -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.sqt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str) -> int:
        if rc < 0:
            raise SqtError("%s failed (%d): %s" % (what, rc, self.lib.sqt_last_error(self.h).decode()))
        return rc

    # ---- the reference seam ----------------------------------------------------------------
    def set_limit(self, limitCorrAv: int):
        self._check(self.lib.sqt_set_limit(self.h, int(limitCorrAv)), "sqt_set_limit")

    def InitSpacetimeCorr(self):
        self._check(self.lib.sqt_begin_file(self.h), "sqt_begin_file")

    begin_file = InitSpacetimeCorr

    def set_types(self, types: np.ndarray):
        """Per-type partials: the species (1 .. n_types) of this shard's atoms in dump order, before the first frame."""
        t = np.ascontiguousarray(types, dtype=np.int32)
        self._check(self.lib.sqt_set_types(self.h, t.ctypes.data, int(t.shape[0])), "sqt_set_types")

    def stage(self, n_local: int) -> np.ndarray:
        """Pinned SoA staging for the next frame as a (3, n) float64 view ((6, n) with -V: x y z vx vy vz)."""
        p = self.lib.sqt_stage_acquire(self.h, int(n_local))
        if not p:
            raise SqtError("sqt_stage_acquire: " + self.lib.sqt_last_error(self.h).decode())
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(self.ncomp, n_local))

    def EvalSpacetimeCorr(self, xyz: np.ndarray, xlo: float, xhi: float, n_total: int | None = None, vel: np.ndarray | None = None,
                          mu: np.ndarray | None = None) -> int:
        """One frame.  ``xyz`` is (n, 3) atoms in dump order (this rank's shard); ``vel`` likewise with -V / -s, ``mu`` with
        -M.  Returns SQT_FLUSH (1) when the reference would have called PrintProcess inside AccumSpacetimeCorr."""
        n = xyz.shape[0]
        st = self.stage(n)
        _fill_stage(st, self.flags, xyz, vel, mu)
        return self.push_staged(st, n, n_total if n_total is not None else n, xlo, xhi)

    push_frame = EvalSpacetimeCorr

    def push_staged(self, st: np.ndarray, n_local: int, n_total: int, xlo: float, xhi: float) -> int:
        return self._check(self.lib.sqt_push_frame(self.h, st.ctypes.data, int(n_local), int(n_total), float(xlo), float(xhi)),
                           "sqt_push_frame")

    def push_device(self, dptr: int, n_local: int, n_total: int, xlo: float, xhi: float) -> int:
        return self._check(self.lib.sqt_push_frame_device(self.h, C.c_void_p(dptr), int(n_local), int(n_total), float(xlo), float(xhi)),
                           "sqt_push_frame_device")

    def fetch(self) -> Results:
        r = Results(self.S, self.T, self.n_types)
        self._check(self.lib.sqt_fetch_results(self.h, C.byref(r.c)), "sqt_fetch_results")
        return r

    def finish(self) -> Results | None:
        """End of run (lmp2sqt.c:327-333): results to print, or None if a mid-run print already happened."""
        r = Results(self.S, self.T, self.n_types)
        rc = self._check(self.lib.sqt_finish(self.h, C.byref(r.c)), "sqt_finish")
        return r if rc == 1 else None

    def PrintProcess(self, outdir: str, results: Results | None = None, date: str | None = None) -> Results:
        r = results if results is not None else self.fetch()
        r.print_process(outdir, self.deck, self.flags, date)
        return r

    # ---- introspection ---------------------------------------------------------------------
    def sync(self):
        self._check(self.lib.sqt_sync(self.h), "sqt_sync")

    @property
    def countCorrAv(self) -> int:
        return self.lib.sqt_count_corr_av(self.h)

    @property
    def number_call_print(self) -> int:
        return self.lib.sqt_number_call_print(self.h)

    def buffer_counts(self):
        return [self.lib.sqt_buffer_count(self.h, nb) for nb in range(self.B)]

    def rho(self) -> np.ndarray:
        out = np.zeros((3, self.S, 2))
        self._check(self.lib.sqt_get_rho(self.h, out.ctypes.data), "sqt_get_rho")
        return out

    def kvel(self) -> np.ndarray:
        out = np.zeros((3, self.S, 2))
        self._check(self.lib.sqt_get_kvel(self.h, out.ctypes.data), "sqt_get_kvel")
        return out

    def kmu(self) -> np.ndarray:
        out = np.zeros((3, self.S, 2))
        self._check(self.lib.sqt_get_kmu(self.h, out.ctypes.data), "sqt_get_kmu")
        return out

    @property
    def kernel_launches(self) -> int:
        return self.lib.sqt_kernel_launches(self.h)

    @property
    def kernel_config(self) -> str:
        return self.lib.sqt_kernel_config(self.h).decode()

    def profile(self, on: bool):
        self.lib.sqt_profile_enable(self.h, int(on))

    def profile_read(self):
        n, ms = C.c_longlong(), C.c_double()
        self._check(self.lib.sqt_profile_read(self.h, C.byref(n), C.byref(ms)), "sqt_profile_read")
        return n.value, ms.value

    def timer_start(self):
        self._check(self.lib.sqt_timer_start(self.h), "sqt_timer_start")

    def timer_stop(self) -> float:
        ms = C.c_double()
        self._check(self.lib.sqt_timer_stop(self.h, C.byref(ms)), "sqt_timer_stop")
        return ms.value

    # ---- multi-GPU -------------------------------------------------------------------------
    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(unique_id, _lib.SQT_COMM_ID_BYTES)
        self._check(self.lib.sqt_comm_init(self.h, buf, rank, nranks), "sqt_comm_init")


class MultiSpacetimeCorr:
    """G GPUs behind one caller thread (include/sqt.h, sqt_multi_*): one pinned frame, fan-out by particle chunk."""

    def __init__(self, deck: Deck, flags: str = "Ft", n_gpus: int = 2, device: int = 0, n_types: int = 0):
        self.lib = _lib.engine_lib()
        self.deck, self.flags = deck, flags
        self.n_types = n_types if n_types > 1 else 0
        ff, fF, ft, fV, fM, fs = _flag_ints(flags)
        p = SqtParams(deck.kVal, deck.deltaT, deck.mass, deck.rVal, deck.limitCorrAv, deck.nCBuffer, deck.nCSpatial,
                      deck.nCTime, deck.nCSkip, ff, fF, ft, device, n_gpus, fV, fM, fs, n_types)
        self.ncomp = _ncomp(flags)
        h = C.c_void_p()
        rc = self.lib.sqt_multi_create(C.byref(p), C.byref(h))
        if rc:
            raise SqtError("sqt_multi_create failed (%d): %s" % (rc, self.lib.sqt_multi_last_error(None).decode()))
        self.h = h
        self.S, self.T = deck.nCSpatial, deck.nCTime

    def close(self):
        if getattr(self, "h", None):
            self.lib.sqt_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str) -> int:
        if rc < 0:
            raise SqtError("%s failed (%d): %s" % (what, rc, self.lib.sqt_multi_last_error(self.h).decode()))
        return rc

    def set_limit(self, v: int):
        self._check(self.lib.sqt_multi_set_limit(self.h, int(v)), "sqt_multi_set_limit")

    def InitSpacetimeCorr(self):
        self._check(self.lib.sqt_multi_begin_file(self.h), "sqt_multi_begin_file")

    def EvalSpacetimeCorr(self, xyz, xlo, xhi, n_total=None, vel=None, mu=None) -> int:
        n = xyz.shape[0]
        p = self.lib.sqt_multi_stage_acquire(self.h, n)
        if not p:
            raise SqtError("sqt_multi_stage_acquire: " + self.lib.sqt_multi_last_error(self.h).decode())
        st = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(self.ncomp, n))
        _fill_stage(st, self.flags, xyz, vel, mu)
        return self._check(self.lib.sqt_multi_push_frame(self.h, st.ctypes.data, n, float(xlo), float(xhi)), "sqt_multi_push_frame")

    def set_types(self, types: np.ndarray):
        t = np.ascontiguousarray(types, dtype=np.int32)
        self._check(self.lib.sqt_multi_set_types(self.h, t.ctypes.data, int(t.shape[0])), "sqt_multi_set_types")

    def fetch(self) -> Results:
        r = Results(self.S, self.T, self.n_types)
        self._check(self.lib.sqt_multi_fetch_results(self.h, C.byref(r.c)), "sqt_multi_fetch_results")
        return r

    def finish(self) -> Results | None:
        r = Results(self.S, self.T, self.n_types)
        rc = self._check(self.lib.sqt_multi_finish(self.h, C.byref(r.c)), "sqt_multi_finish")
        return r if rc == 1 else None

    @property
    def n_gpus(self) -> int:
        return self.lib.sqt_multi_n_gpus(self.h)

    @property
    def kernel_config(self) -> str:
        return self.lib.sqt_kernel_config(self.lib.sqt_multi_handle(self.h, 0)).decode()

    @property
    def kernel_launches(self) -> int:
        return self.lib.sqt_multi_kernel_launches(self.h)

    @property
    def countCorrAv(self) -> int:
        return self.lib.sqt_count_corr_av(self.lib.sqt_multi_handle(self.h, 0))

    @property
    def number_call_print(self) -> int:
        return self.lib.sqt_number_call_print(self.lib.sqt_multi_handle(self.h, 0))


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(_lib.SQT_COMM_ID_BYTES)
    rc = _lib.engine_lib().sqt_comm_unique_id(buf)
    if rc:
        raise SqtError("sqt_comm_unique_id failed: " + _lib.engine_lib().sqt_last_error(None).decode())
    return buf.raw


class SynthSource:
    """On-device synthetic frame source (include/sqt.h, sqt_synth_*)."""

    def __init__(self, n_cells: int, first_atom: int, n_local: int, seed: int = 20261019, device: int = 0):
        self.lib = _lib.engine_lib()
        h = C.c_void_p()
        rc = self.lib.sqt_synth_create(device, n_cells, first_atom, n_local, seed, C.byref(h))
        if rc:
            raise SqtError("sqt_synth_create failed (%d)" % rc)
        self.h = h
        self.n_local = n_local
        self.n_total = 4 * n_cells ** 3
        self.L = self.lib.sqt_synth_box_length(h)

    def next(self) -> int:
        p = self.lib.sqt_synth_next(self.h)
        if not p:
            raise SqtError("sqt_synth_next failed")
        return p

    def to_host(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((3, self.n_local))
        rc = self.lib.sqt_synth_copy_to_host(self.h, out.ctypes.data)
        if rc:
            raise SqtError("sqt_synth_copy_to_host failed")
        return out

    def to_device(self, dptr: int) -> None:
        rc = self.lib.sqt_synth_copy_to_device(self.h, C.c_void_p(dptr))
        if rc:
            raise SqtError("sqt_synth_copy_to_device failed")

    def close(self):
        if getattr(self, "h", None):
            self.lib.sqt_synth_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def measure_dfma_peak(device: int = 0):
    tf, mhz = C.c_double(), C.c_double()
    rc = _lib.engine_lib().sqt_measure_dfma_peak(device, C.byref(tf), C.byref(mhz))
    if rc:
        raise SqtError("sqt_measure_dfma_peak failed (%d)" % rc)
    return tf.value, mhz.value


def shard_bounds(n_total: int, rank: int, nranks: int) -> tuple[int, int]:
    """Particle shard [lo, hi) of ``rank``: contiguous in dump order."""
    lo = (n_total * rank) // nranks
    hi = (n_total * (rank + 1)) // nranks
    return lo, hi


def run_trajectories(trajs, deck: Deck, flags: str = "Ft", device: int = 0, rank: int = 0, nranks: int = 1,
                     unique_id: bytes | None = None, on_flush=None, n_gpus: int = 0, n_types: int = 0):
    """Drive the engine over a list of trajectories (= input files) exactly like the reference's main
    (lmp2sqt.c:200-333), including its limitCorrAv bookkeeping.  Returns (printed Results | None, engine).
    ``n_gpus`` > 0 uses the in-process multi-GPU fan-out (sqt_multi_*) instead of one handle."""
    B = max(1, min(deck.nCBuffer, deck.nCTime))
    limit = deck.limitCorrAv
    use = []
    for t in trajs:
        n_snap = t.n_frames
        ok = n_snap >= 5 and int((n_snap - deck.nCTime) / (deck.nCTime // B)) >= 1  # lmp2sqt.c:221-237
        use.append(ok)
        if ok:
            limit += (n_snap - deck.nCTime) // (deck.nCTime // B)                    # lmp2sqt.c:244
    if n_gpus > 0:
        eng = MultiSpacetimeCorr(deck, flags, n_gpus, device, n_types)
    else:
        eng = SpacetimeCorr(deck, flags, device, n_types)
        if nranks > 1:
            eng.comm_init(unique_id, rank, nranks)
    eng.set_limit(limit)
    if n_types > 1:  # the dump's `type` column; atoms keep their order, so the first used file's species hold for the run
        t0 = next(t for t, u in zip(trajs, use) if u)
        lo, hi = (0, t0.n_atoms) if n_gpus > 0 else shard_bounds(t0.n_atoms, rank, nranks)
        eng.set_types(t0.types[lo:hi])
    printed = None
    with_v, with_m = ("V" in flags or "s" in flags), "M" in flags
    for t, u in zip(trajs, use):
        if not u:
            continue
        eng.InitSpacetimeCorr()
        lo, hi = shard_bounds(t.n_atoms, rank, nranks)
        for f in range(t.n_frames):
            rc = eng.EvalSpacetimeCorr(t.positions[f, lo:hi], t.box[f, 0, 0], t.box[f, 0, 1], n_total=t.n_atoms,
                                       vel=t.velocities[f, lo:hi] if with_v else None, mu=t.dipoles[f, lo:hi] if with_m else None)
            if rc == SQT_FLUSH:
                printed = eng.fetch()
                if on_flush:
                    on_flush(printed)
    fin = eng.finish()
    if fin is not None:
        printed = fin
    return printed, eng
