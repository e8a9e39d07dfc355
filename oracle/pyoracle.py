"""TEST INFRASTRUCTURE ONLY: ctypes access to the two CPU checkers under ``oracle/``.

* :class:`OraclePort`  -- ``oracle/libsqt_oracle.so``, our plain-C restatement (sqt_oracle.c).
* :class:`ReferenceLib` -- ``oracle/_ref/liblmp2sqt_ref.so``, the UNMODIFIED reference compiled
  as a shared library by ``oracle/Makefile`` (``main`` renamed on the command line), driven
  in-process exactly the way the reference's ``main`` drives it (lmp2sqt.c:117-336) so the raw
  fp64 globals can be read before ``prePrintProcess`` rounds them into ``%8.4g`` text.
  The reference keeps all state in file-scope globals, so use one *process* per run:
  :func:`run_reference` does that for you.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl reference``
legs may import this module.  The product package ``lmp2sqt_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import pickle
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libsqt_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "liblmp2sqt_ref.so")
REF_BIN = os.path.join(HERE, "_ref", "lmp2sqt_ref.out")
REF_ANALYSIS_BIN = os.path.join(HERE, "_ref", "analysis_binary_ref.out")
REF_ANALYSIS_TEXT_BIN = os.path.join(HERE, "_ref", "analysis_ref.out")   # analysis.c (= AverageFqt2FqOrFt.c), reads output.txt

NAMES2D = ["avF", "StdDevF", "avF_s", "StdDevF_s", "avF_d", "StdDevF_d", "avDrTable"]
NAMES1D = ["rrMSDAv", "rrMQDAv", "rrMSDCMAv"]
# -V (flag_velocity) adds these (oracle array indices 7..9 and 3..4)
NAMES2D_V = ["avC2_v_rho", "avC2_rho_v", "avC2_v_v"]
NAMES1D_V = ["rrCvvAv", "rrCvcmvcmAv"]
# -M (flag_magnet): oracle 2-D indices 10..14 (10, 11 only with -V as well) and 1-D index 5 (rrCmmAv: allocated, zeroed, printed,
# never added to -- lmp2sqt.c:420-432 has no line for it)
NAMES2D_VM = ["avC2_v_mu", "avC2_mu_v"]
NAMES2D_M = ["avC2_mu_rho", "avC2_rho_mu", "avC2_mu_mu"]
NAMES1D_M = ["rrCmmAv"]
# -s (flag_s): oracle 1-D indices 6..17 raw (+ 18, 19 in the printed set): rrMSR1_R_Av [T][3], rrMSR2_VR_Av [T][9]
NAMES1D_S = ["rrMSR1_R_Av", "rrMSR2_VR_Av"]


def names_for(flags):
    """(2-D names with their oracle index, 1-D names with index) of a flag set."""
    n2 = list(enumerate(NAMES2D))
    n1 = list(enumerate(NAMES1D))
    if "V" in flags:
        n2 += [(7 + i, n) for i, n in enumerate(NAMES2D_V)]
        n1 += [(3 + i, n) for i, n in enumerate(NAMES1D_V)]
    if "V" in flags and "M" in flags:
        n2 += [(10 + i, n) for i, n in enumerate(NAMES2D_VM)]
    if "M" in flags:
        n2 += [(12 + i, n) for i, n in enumerate(NAMES2D_M)]
        n1 += [(5, "rrCmmAv")]
    return n2, n1


def build(force: bool = False) -> None:
    """Compile the restatement (always possible) and, where /root/reference exists, the reference."""
    if force or not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(os.path.join(HERE, "sqt_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    if os.path.exists("/root/reference/lmp2sqt.c") and (force or not os.path.exists(REF_SO) or not os.path.exists(REF_ANALYSIS_TEXT_BIN)):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def have_reference() -> bool:
    return os.path.exists(REF_SO) and os.path.exists(REF_BIN)


class _Params(C.Structure):
    _fields_ = [("kVal", C.c_double), ("deltaT", C.c_double), ("mass", C.c_double), ("rVal", C.c_double),
                ("limitCorrAv", C.c_int), ("nCBuffer", C.c_int), ("nCSpatial", C.c_int), ("nCTime", C.c_int),
                ("nCSkip", C.c_int), ("flag_f", C.c_int), ("flag_F", C.c_int), ("flag_t", C.c_int), ("flag_V", C.c_int),
                ("flag_M", C.c_int), ("flag_s", C.c_int)]


class OraclePort:
    """The plain-C restatement.  Method names follow the reference's call seam."""

    def __init__(self, nCSpatial, nCTime, nCBuffer, nCSkip=0, rVal=0.2, limitCorrAv=0, deltaT=0.06,
                 kVal=0.0981748, mass=1.0, flag_f=True, flag_F=True, flag_t=True, flag_V=False, flag_M=False, flag_s=False):
        build()
        L = C.CDLL(PORT_SO)
        L.sqt_oracle_create.restype = C.c_void_p
        L.sqt_oracle_create.argtypes = [C.POINTER(_Params)]
        for fn in ("sqt_oracle_destroy", "sqt_oracle_begin_file"):
            getattr(L, fn).argtypes = [C.c_void_p]
            getattr(L, fn).restype = None
        L.sqt_oracle_push_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.sqt_oracle_push_frame_v.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.sqt_oracle_push_frame_vm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.sqt_oracle_finish.argtypes = [C.c_void_p]
        L.sqt_oracle_set_limit.argtypes = [C.c_void_p, C.c_int]
        L.sqt_oracle_set_limit.restype = None
        for fn in ("countCorrAv", "printed_count", "number_call_print", "limitCorrAv", "nPtls"):
            f = getattr(L, "sqt_oracle_" + fn)
            f.argtypes = [C.c_void_p]
            f.restype = C.c_int
        L.sqt_oracle_buffer_count.argtypes = [C.c_void_p, C.c_int]
        L.sqt_oracle_kVal.argtypes = [C.c_void_p]
        L.sqt_oracle_kVal.restype = C.c_double
        for fn in ("raw2d", "printed2d", "raw1d", "printed1d", "printed_raw2d", "printed_raw1d"):
            f = getattr(L, "sqt_oracle_" + fn)
            f.argtypes = [C.c_void_p, C.c_int]
            f.restype = C.POINTER(C.c_double)
        for fn in ("first_vol", "vol"):
            f = getattr(L, "sqt_oracle_" + fn)
            f.argtypes = [C.c_void_p]
            f.restype = C.c_double
        L.sqt_oracle_first_limit.argtypes = [C.c_void_p]
        for fn in ("factorDr", "rho"):
            f = getattr(L, "sqt_oracle_" + fn)
            f.argtypes = [C.c_void_p]
            f.restype = C.POINTER(C.c_double)
        self.L = L
        self.S, self.T = nCSpatial, nCTime
        self.B = max(1, min(nCBuffer, nCTime))
        p = _Params(kVal, deltaT, mass, rVal, limitCorrAv, nCBuffer, nCSpatial, nCTime, nCSkip,
                    int(flag_f), int(flag_F), int(flag_t), int(flag_V), int(flag_M), int(flag_s))
        self.h = L.sqt_oracle_create(C.byref(p))
        self.flag_V, self.flag_M, self.flag_s = bool(flag_V), bool(flag_M), bool(flag_s)
        self.names2d, self.names1d = names_for(("V" if flag_V else "") + ("M" if flag_M else ""))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.sqt_oracle_destroy(self.h)
            self.h = None

    def begin_file(self):
        self.L.sqt_oracle_begin_file(self.h)

    def push_frame(self, xyz: np.ndarray, xlo: float, xhi: float, vel: np.ndarray | None = None, mu: np.ndarray | None = None) -> int:
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        if mu is not None:
            mu = np.ascontiguousarray(mu, dtype=np.float64)
            vel = None if vel is None else np.ascontiguousarray(vel, dtype=np.float64)
            rc = self.L.sqt_oracle_push_frame_vm(self.h, xyz.ctypes.data, None if vel is None else vel.ctypes.data, mu.ctypes.data,
                                                 xyz.shape[0], float(xlo), float(xhi))
        elif vel is not None:
            vel = np.ascontiguousarray(vel, dtype=np.float64)
            rc = self.L.sqt_oracle_push_frame_v(self.h, xyz.ctypes.data, vel.ctypes.data, xyz.shape[0], float(xlo), float(xhi))
        else:
            rc = self.L.sqt_oracle_push_frame(self.h, xyz.ctypes.data, xyz.shape[0], float(xlo), float(xhi))
        if rc < 0:
            raise RuntimeError("oracle push_frame failed rc=%d" % rc)
        return rc

    def finish(self) -> int:
        return self.L.sqt_oracle_finish(self.h)

    def set_limit(self, v):
        self.L.sqt_oracle_set_limit(self.h, int(v))

    @property
    def countCorrAv(self):
        return self.L.sqt_oracle_countCorrAv(self.h)

    @property
    def printed_count(self):
        return self.L.sqt_oracle_printed_count(self.h)

    @property
    def kVal(self):
        return self.L.sqt_oracle_kVal(self.h)

    def buffer_counts(self):
        return [self.L.sqt_oracle_buffer_count(self.h, nb) for nb in range(self.B)]

    def _arr(self, fn, which, shape):
        p = getattr(self.L, "sqt_oracle_" + fn)(self.h, which)
        return np.ctypeslib.as_array(p, shape=shape).copy()

    def _set(self, kind) -> dict:
        d = {n: self._arr(kind + "2d", i, (self.S, self.T)) for i, n in self.names2d}
        d.update({n: self._arr(kind + "1d", i, (self.T,)) for i, n in self.names1d})
        if self.flag_s:   # the reference's VecR3 / Rank2R3 arrays: [T][3] and [T][9] (xx xy xz yx yy yz zx zy zz)
            d["rrMSR1_R_Av"] = np.stack([self._arr(kind + "1d", 6 + j, (self.T,)) for j in range(3)], axis=1)
            d["rrMSR2_VR_Av"] = np.stack([self._arr(kind + "1d", 9 + j, (self.T,)) for j in range(9)], axis=1)
            if kind == "printed":
                d["rrMSR2_VR_Av_dig"] = self._arr("printed1d", 18, (self.T,))
                d["rrMSR2_VR_Av_offdig"] = self._arr("printed1d", 19, (self.T,))
        return d

    def raw(self) -> dict:
        d = self._set("raw")
        d["countCorrAv"] = self.countCorrAv
        return d

    def printed(self) -> dict:
        d = self._set("printed")
        d["countCorrAv"] = self.printed_count
        d["kVal"] = self.kVal
        return d

    def printed_raw(self) -> dict:
        """Raw sums at the moment of the last print + the scalars the host normalisation needs."""
        d = self._set("printed_raw")
        d.update(countCorrAv=self.printed_count, kVal=self.kVal, nPtls=self.L.sqt_oracle_nPtls(self.h),
                 first_Vol=self.L.sqt_oracle_first_vol(self.h), first_limitCorrAv=self.L.sqt_oracle_first_limit(self.h),
                 g_Vol=self.L.sqt_oracle_vol(self.h), limitCorrAv=self.L.sqt_oracle_limitCorrAv(self.h))
        return d

    def rho(self):
        return np.ctypeslib.as_array(self.L.sqt_oracle_rho(self.h), shape=(3, self.S, 2)).copy()

    def factorDr(self):
        return np.ctypeslib.as_array(self.L.sqt_oracle_factorDr(self.h), shape=(self.S,)).copy()


def run_port(traj, nCSpatial, nCTime, nCBuffer, nCSkip=0, rVal=0.2, limitCorrAv=0, flags="Ft",
             files=None, finish=True):
    """Drive the restatement over a trajectory (or a list of trajectories = several dump files)
    the way the reference's ``main`` would, including its limitCorrAv bookkeeping
    (lmp2sqt.c:221-247).  Returns (printed dict | None, raw dict, per-frame log)."""
    trajs = files if files is not None else [traj]
    flag_F = "F" in flags
    flag_f = flag_F or "f" in flags
    flag_t = "t" in flags
    flag_V = "V" in flags
    flag_M = "M" in flags
    flag_s = "s" in flags
    flag_t = flag_t or flag_s   # getopt: -s switches -t on as well (lmp2sqt.c:156-160)
    o = OraclePort(nCSpatial, nCTime, nCBuffer, nCSkip, rVal, 0, flag_f=flag_f, flag_F=flag_F, flag_t=flag_t, flag_V=flag_V,
                   flag_M=flag_M, flag_s=flag_s)
    limit = limitCorrAv
    B = min(nCBuffer, nCTime)
    use = []
    for t in trajs:
        n_snap = t.n_frames
        if n_snap < 5 or int((n_snap - nCTime) / (nCTime // B)) < 1:  # C int division, lmp2sqt.c:228
            use.append(False)
            continue
        limit += (n_snap - nCTime) // (nCTime // B)  # lmp2sqt.c:244
        use.append(True)
    o.set_limit(limit)
    log = []
    printed = None
    for t, u in zip(trajs, use):
        if not u:
            continue
        o.begin_file()
        for f in range(t.n_frames):
            rc = o.push_frame(t.positions[f], t.box[f, 0, 0], t.box[f, 0, 1], vel=t.velocities[f] if (flag_V or flag_s) else None,
                              mu=t.dipoles[f] if flag_M else None)
            log.append((o.countCorrAv, tuple(o.buffer_counts())))
            if rc == 1:
                printed = o.printed()
                printed["_raw_at_print"] = o.printed_raw()
    raw = o.raw()
    if finish and o.finish():
        printed = o.printed()
    if printed is not None:
        printed["_raw_at_print"] = o.printed_raw()
    return printed, raw, log


# ----------------------------------------------------------------------------------------------
# The unmodified reference, in-process.
# ----------------------------------------------------------------------------------------------
ATOM_DTYPE = np.dtype([("id", "i4"), ("type", "i4"), ("x", "f8"), ("y", "f8"), ("z", "f8"),
                       ("mux", "f8"), ("muy", "f8"), ("muz", "f8"), ("vx", "f8"), ("vy", "f8"),
                       ("vz", "f8"), ("mu1", "f8"), ("atomType", "i4")], align=True)  # snapshot.h:95-110
assert ATOM_DTYPE.itemsize == 96


class _Box3(C.Structure):  # snapshot.h:118-126
    _fields_ = [("xlow", C.c_double), ("xhigh", C.c_double), ("ylow", C.c_double), ("yhigh", C.c_double),
                ("zlow", C.c_double), ("zhigh", C.c_double), ("pbcTYPE", C.c_int)]


class _Snapshot(C.Structure):  # snapshot.h:132-138
    _fields_ = [("timestep", C.c_long), ("NumAtoms", C.c_int), ("Box", _Box3), ("atoms", C.c_void_p),
                ("snapFlag", C.c_int)]


class _MakeSqtClass(C.Structure):  # lmp2sqt.h:156-162
    _fields_ = [("tBuf", C.c_void_p), ("flag_alloc", C.c_int), ("flag_alloc_more", C.c_int),
                ("snap", C.POINTER(_Snapshot)), ("nSkip", C.c_int)]


class ReferenceLib:
    """Drives the reference's own functions.  ONE instance per process (file-scope globals)."""

    def __init__(self, nCSpatial, nCTime, nCBuffer, nCSkip=0, rVal=0.2, limitCorrAv=0, deltaT=0.06,
                 kVal=0.0981748, mass=1.0, flags="Ft"):
        L = C.CDLL(REF_SO)
        self.L = L
        g = lambda name, t: t.in_dll(L, name)
        self._g = g
        g("kVal", C.c_double).value = kVal
        g("deltaT", C.c_double).value = deltaT
        g("mass", C.c_double).value = mass
        g("rVal", C.c_double).value = rVal
        g("nCSpatial", C.c_int).value = nCSpatial
        g("nCTime", C.c_int).value = nCTime
        g("nCBuffer", C.c_int).value = max(1, min(nCBuffer, nCTime))  # lmp2sqt.c:2105, 240-242
        g("nCSkip", C.c_int).value = nCSkip
        g("limitCorrAv", C.c_int).value = limitCorrAv
        flag_F = "F" in flags
        g("flag_F", C.c_bool).value = flag_F
        g("flag_f", C.c_bool).value = flag_F or ("f" in flags)
        self.flag_V, self.flag_M, self.flag_s = "V" in flags, "M" in flags, "s" in flags
        g("flag_t", C.c_bool).value = ("t" in flags) or self.flag_s   # getopt, lmp2sqt.c:156-160
        g("flag_velocity", C.c_bool).value = self.flag_V
        g("flag_magnet", C.c_bool).value = self.flag_M
        g("flag_s", C.c_bool).value = self.flag_s
        self.flags = flags
        self.S, self.T, self.B = nCSpatial, nCTime, max(1, min(nCBuffer, nCTime))
        self.cl = _MakeSqtClass()
        C.c_void_p.in_dll(L, "classSqt").value = C.addressof(self.cl)
        for fn in ("InitSpacetimeCorr", "EvalSpacetimeCorr", "PrintProcess"):
            getattr(L, fn).argtypes = [C.POINTER(_MakeSqtClass)]
            getattr(L, fn).restype = None
        L.prePrintProcess.restype = None
        self._keep = None

    def set_limit(self, v):
        self._g("limitCorrAv", C.c_int).value = int(v)

    def begin_file(self):
        self.L.InitSpacetimeCorr(C.byref(self.cl))

    def push_frame(self, xyz: np.ndarray, xlo, xhi, box_yz=None, vel=None, mu=None) -> float:
        """EvalSpacetimeCorr on a Snapshot built from ``xyz`` (and ``vel``); returns the call's wall seconds."""
        n = xyz.shape[0]
        atoms = np.zeros(n, dtype=ATOM_DTYPE)
        atoms["id"] = np.arange(1, n + 1)
        atoms["type"] = 1
        atoms["x"], atoms["y"], atoms["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
        if vel is not None:
            atoms["vx"], atoms["vy"], atoms["vz"] = vel[:, 0], vel[:, 1], vel[:, 2]
        if mu is not None:
            atoms["mux"], atoms["muy"], atoms["muz"] = mu[:, 0], mu[:, 1], mu[:, 2]
        snap = _Snapshot()
        snap.timestep = 0
        snap.NumAtoms = n
        ylo, yhi, zlo, zhi = box_yz if box_yz is not None else (xlo, xhi, xlo, xhi)
        snap.Box = _Box3(xlo, xhi, ylo, yhi, zlo, zhi, 7)
        snap.atoms = atoms.ctypes.data
        self._keep = (atoms, snap)
        self.cl.snap = C.pointer(snap)
        t0 = time.perf_counter()
        self.L.EvalSpacetimeCorr(C.byref(self.cl))
        return time.perf_counter() - t0

    @property
    def countCorrAv(self):
        return self._g("countCorrAv", C.c_int).value

    @property
    def number_call_print(self):
        return self._g("Number_call_Print", C.c_int).value

    def _a2(self, name):
        pp = C.POINTER(C.POINTER(C.c_double)).in_dll(self.L, name)
        if not pp:
            return np.zeros((self.S, self.T))
        return np.ctypeslib.as_array(pp[0], shape=(self.S, self.T)).copy()  # AllocMem2: one contiguous block

    def _a1(self, name):
        p = C.POINTER(C.c_double).in_dll(self.L, name)
        if not p:
            return np.zeros(self.T)
        return np.ctypeslib.as_array(p, shape=(self.T,)).copy()

    def arrays(self) -> dict:
        m = {"avF": "avF_qq2", "StdDevF": "StdDevF_qq2", "avF_s": "avF_s_qq2", "StdDevF_s": "StdDevF_s_qq2",
             "avF_d": "avF_d_qq2", "StdDevF_d": "StdDevF_d_qq2", "avDrTable": "avDrTable"}
        n2, n1 = names_for(self.flags)
        m.update({k: k for _, k in n2 if k not in NAMES2D})
        d = {k: self._a2(v) for k, v in m.items()}
        for _, n in n1:
            d[n] = self._a1(n)
        if self.flag_s:   # VecR3[T], Rank2R3[T]
            for name, w in (("rrMSR1_R_Av", 3), ("rrMSR2_VR_Av", 9), ("rrMSR2_VR_Av_dig", 1), ("rrMSR2_VR_Av_offdig", 1)):
                p = C.POINTER(C.c_double).in_dll(self.L, name)
                d[name] = np.ctypeslib.as_array(p, shape=(self.T, w) if w > 1 else (self.T,)).copy()
        d["countCorrAv"] = self.countCorrAv
        d["kVal"] = self._g("kVal", C.c_double).value
        return d

    def pre_print(self):
        """prePrintProcess (lmp2sqt.c:576-675): normalise the sums in place."""
        self.L.prePrintProcess()


def _ref_worker(job_path: str) -> None:
    with open(job_path, "rb") as f:
        job = pickle.load(f)
    os.chdir(job["cwd"])
    ref = ReferenceLib(**job["params"])
    ref.set_limit(job["limit"])
    times, log = [], []
    for pos, box, vel, mu in job["files"]:
        ref.begin_file()
        for fi in range(pos.shape[0]):
            dt = ref.push_frame(pos[fi], box[fi, 0, 0], box[fi, 0, 1], vel=None if vel is None else vel[fi], mu=None if mu is None else mu[fi])
            times.append(dt)
            log.append(ref.countCorrAv)
    raw = ref.arrays()
    out = {"raw": raw, "times": times, "log": log, "n_print": ref.number_call_print}
    if job.get("print_files") and raw["countCorrAv"] > 0:
        # the reference's own PrintProcess (lmp2sqt.c:338-367) in the scratch cwd: every file the flag set produces, with the date
        # suffix "GOLD" (datetime_data is main's job, :127-132); its C2_* streams are never closed, hence the fflush(NULL)
        C.memmove(C.addressof((C.c_char * 100).in_dll(ref.L, "datetime_data")), b"GOLD\0", 5)
        ref.L.PrintProcess(C.byref(ref.cl))
        C.CDLL(None).fflush(None)
        out["files"] = {fn: open(fn, "rb").read() for fn in sorted(os.listdir(".")) if fn.endswith(".GOLD") or fn.startswith("output.")}
    elif job.get("normalise", True) and raw["countCorrAv"] > 0:
        ref.pre_print()
        out["printed"] = ref.arrays()
        out["printed"]["countCorrAv"] = raw["countCorrAv"]
    with open(job["out"], "wb") as f:
        pickle.dump(out, f)


def run_reference(trajs, nCSpatial, nCTime, nCBuffer, nCSkip=0, rVal=0.2, flags="Ft", threads=1,
                  limit=10**9, normalise=True, print_files=False) -> dict:
    """Run the unmodified reference in a fresh process over ``trajs`` (list of Trajectory = input
    files).  ``limit`` is the in-memory limitCorrAv (default: never flush mid-run, so that the raw
    sums of the whole run can be read).  OMP_NUM_THREADS=``threads`` (1 for anything MSD-like: the
    reference's loop at lmp2sqt.c:1377 races)."""
    if not have_reference():
        raise RuntimeError("oracle/_ref not built (run `make -C oracle ref` where /root/reference exists)")
    with tempfile.TemporaryDirectory() as td:
        job = {"cwd": td, "out": os.path.join(td, "out.pkl"), "limit": int(limit), "normalise": normalise, "print_files": print_files,
               "params": dict(nCSpatial=nCSpatial, nCTime=nCTime, nCBuffer=nCBuffer, nCSkip=nCSkip, rVal=rVal, flags=flags),
               "files": [(t.positions, t.box, t.velocities if ("V" in flags or "s" in flags) else None,
                          t.dipoles if "M" in flags else None) for t in trajs]}
        jp = os.path.join(td, "job.pkl")
        with open(jp, "wb") as f:
            pickle.dump(job, f)
        # MALLOC_PERTURB_=255 makes glibc hand out zero-filled blocks: the reference never
        # initialises TBuf.rrMSDCM (lmp2sqt.c:1325-1333 vs :1450); zero is what a fresh heap gives its CLI.
        env = dict(os.environ, OMP_NUM_THREADS=str(threads), MALLOC_PERTURB_="255",
                   GLIBC_TUNABLES="glibc.malloc.tcache_count=0")  # tcache hits skip the perturb fill
        code = "import sys; sys.path.insert(0, %r); import pyoracle; pyoracle._ref_worker(%r)" % (HERE, jp)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference worker failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
        with open(job["out"], "rb") as f:
            return pickle.load(f)


def run_reference_cli(workdir, deck_name, dump_names, flags=("-F", "-t"), threads=1, zero_heap=False) -> subprocess.CompletedProcess:
    """Run the unmodified reference *binary* in ``workdir`` (it writes its outputs and rewrites the deck there).
    ``zero_heap``: glibc hands out zero-filled blocks (see run_reference) -- the reference never initialises TBuf.rrMSDCM, and
    on some runs the MSDCM column of its Dt file is whatever the heap held."""
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    if zero_heap:
        env.update(MALLOC_PERTURB_="255", GLIBC_TUNABLES="glibc.malloc.tcache_count=0")
    return subprocess.run([REF_BIN, *flags, "-i", deck_name, *dump_names], cwd=workdir, env=env,
                          capture_output=True, text=True)


def parse_output_binary(path, nCSpatial, nCTime) -> dict:
    """Parse the reference's output.binary (layout: lmp2sqt.c:768-864).  Arrays come back [nCTime][nCSpatial]."""
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    off = 0

    def take(fmt, n=1):
        nonlocal off
        a = np.frombuffer(data, dtype=fmt, count=n, offset=off)
        off += a.nbytes
        return a

    hl = int(take("<i4")[0])
    out["magic"] = data[off:off + hl].decode()
    off += hl
    ntypes = int(take("<i4")[0])
    out["sets"] = []
    while off < len(data):
        for _ in range(ntypes):
            ln = int(take("<i4")[0])
            name = data[off:off + ln].decode()
            off += ln
            kVal, dt, rVal = take("<f8", 3)
            yes = int(take("<i4")[0])
            arr = take("<f8", nCSpatial * nCTime).reshape(nCTime, nCSpatial).copy() if yes else None
            out[name] = arr
            out["kVal"], out["deltaT"], out["rVal"] = float(kVal), float(dt), float(rVal)
        break
    out["size"] = len(data)
    out["consumed"] = off
    return out


if __name__ == "__main__":
    build(force=True)
    print(json.dumps({"port": os.path.exists(PORT_SO), "reference": have_reference()}))


# ----------------------------------------------------------------------------------------------
# CPU baseline for bench.py: the unmodified reference timed on a bounded sample of a workload.
# ----------------------------------------------------------------------------------------------
def _bench_worker(job_path: str) -> None:
    with open(job_path, "rb") as f:
        job = pickle.load(f)
    os.chdir(job["cwd"])
    S, T, B, n = job["S"], job["T"], job["B"], job["n_atoms"]
    ref = ReferenceLib(S, T, B, 0, 0.2, flags="Ft")
    ref.set_limit(10**9)
    rng = np.random.Generator(np.random.PCG64(job["seed"]))
    L = float((n / 0.8442) ** (1.0 / 3.0))
    pos = rng.uniform(0.0, L, size=(n, 3))
    ref.begin_file()
    times = []
    total = job["fill"] + job["warmup"] + job["steps"]
    for f in range(total):
        pos = pos + rng.normal(0.0, 0.077, size=(n, 3))
        times.append(ref.push_frame(pos, 0.0, L))
    with open(job["out"], "wb") as f:
        pickle.dump({"times": times, "countCorrAv": ref.countCorrAv}, f)


def bench_reference(n_atoms, S, T, B, fill, warmup, steps, threads, seed=1, timeout=1500) -> dict:
    """Time the reference's own EvalSpacetimeCorr (lmp2sqt.c:1599) frame by frame on ``n_atoms`` random-walk
    atoms: ``fill`` frames to fill the origin ring (untimed), ``warmup`` + ``steps`` steady-state frames.
    OMP_NUM_THREADS=``threads``.  Returns per-frame wall seconds of the timed steps."""
    if not have_reference():
        raise RuntimeError("oracle/_ref not built")
    with tempfile.TemporaryDirectory() as td:
        job = dict(cwd=td, out=os.path.join(td, "o.pkl"), S=S, T=T, B=B, n_atoms=n_atoms, fill=fill, warmup=warmup,
                   steps=steps, seed=seed)
        jp = os.path.join(td, "job.pkl")
        with open(jp, "wb") as f:
            pickle.dump(job, f)
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        code = "import sys; sys.path.insert(0, %r); import pyoracle; pyoracle._bench_worker(%r)" % (HERE, jp)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError("reference bench worker failed:\n" + r.stdout[-1500:] + r.stderr[-1500:])
        with open(job["out"], "rb") as f:
            out = pickle.load(f)
    t = out["times"]
    out["fill_s"] = float(sum(t[:fill]))
    out["step_times"] = t[fill + warmup:]
    return out
