"""ctypes bindings of the two native libraries (include/sqt.h, include/sqt_host.h).

There is deliberately no fallback: if ``libsqt_b200.so`` is missing the import of the engine
fails loudly -- the product path is the CUDA path or nothing.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ENGINE_SO = os.path.join(PKG_DIR, "libsqt_b200.so")
HOST_SO = os.path.join(PKG_DIR, "libsqt_host.so")
CLI_BIN = os.path.join(PKG_DIR, "bin", "lmp2sqt_b200.out")
ANALYSIS_BIN = os.path.join(PKG_DIR, "bin", "analysis_binary_b200.out")
ANALYSIS_TEXT_BIN = os.path.join(PKG_DIR, "bin", "analysis_b200.out")

SQT_OK, SQT_FLUSH = 0, 1
SQT_COMM_ID_BYTES = 128


class SqtParams(C.Structure):
    _fields_ = [("kVal", C.c_double), ("deltaT", C.c_double), ("mass", C.c_double), ("rVal", C.c_double),
                ("limitCorrAv", C.c_int), ("nCBuffer", C.c_int), ("nCSpatial", C.c_int), ("nCTime", C.c_int),
                ("nCSkip", C.c_int), ("flag_f", C.c_int), ("flag_F", C.c_int), ("flag_t", C.c_int),
                ("device", C.c_int), ("n_gpus", C.c_int), ("flag_V", C.c_int), ("flag_M", C.c_int), ("flag_s", C.c_int),
                ("n_types", C.c_int), ("reserved", C.c_int * 2)]


class SqtFlags(C.Structure):
    _fields_ = [("f", C.c_int), ("F", C.c_int), ("t", C.c_int), ("V", C.c_int), ("M", C.c_int), ("s", C.c_int)]


_PD = C.POINTER(C.c_double)


class SqtResults(C.Structure):
    _fields_ = [("nCSpatial", C.c_int), ("nCTime", C.c_int), ("countCorrAv", C.c_int), ("limitCorrAv", C.c_int),
                ("nPtls", C.c_longlong), ("g_Vol", C.c_double), ("kVal", C.c_double), ("first_Vol", C.c_double),
                ("first_limitCorrAv", C.c_int), ("number_call_print", C.c_int),
                ("avF_qq2", _PD), ("StdDevF_qq2", _PD), ("avF_s_qq2", _PD), ("StdDevF_s_qq2", _PD),
                ("avF_d_qq2", _PD), ("StdDevF_d_qq2", _PD), ("avDrTable", _PD),
                ("rrMSDAv", _PD), ("rrMQDAv", _PD), ("rrMSDCMAv", _PD),
                ("avC2_v_rho", _PD), ("avC2_rho_v", _PD), ("avC2_v_v", _PD), ("rrCvvAv", _PD), ("rrCvcmvcmAv", _PD),
                ("avC2_v_mu", _PD), ("avC2_mu_v", _PD), ("avC2_mu_rho", _PD), ("avC2_rho_mu", _PD), ("avC2_mu_mu", _PD),
                ("rrCmmAv", _PD), ("rrMSR1_R_Av", _PD), ("rrMSR2_VR_Av", _PD),
                ("n_types", C.c_int), ("nPtls_a", C.c_longlong * 8),
                ("avF_ab", _PD), ("avF_s_a", _PD), ("avDrTable_a", _PD), ("rrMSD_a", _PD), ("rrMQD_a", _PD)]


class SqtDeck(C.Structure):
    _fields_ = [("kVal", C.c_double), ("deltaT", C.c_double), ("mass", C.c_double), ("rVal", C.c_double),
                ("limitCorrAv", C.c_int), ("nCBuffer", C.c_int), ("nCSpatial", C.c_int), ("nCTime", C.c_int),
                ("nCSkip", C.c_int), ("status", C.c_int * 9)]


class SqtFrameHeader(C.Structure):
    _fields_ = [("timestep", C.c_long), ("n_atoms", C.c_int), ("xlo", C.c_double), ("xhi", C.c_double),
                ("ylo", C.c_double), ("yhi", C.c_double), ("zlo", C.c_double), ("zhi", C.c_double),
                ("pbcType", C.c_int), ("style", C.c_int)]


class SqtAnalysis(C.Structure):
    _fields_ = [("nCSpatial", C.c_int), ("nCTime", C.c_int), ("nData", C.c_int),
                ("deck_deltaT", C.c_double), ("deltaT", C.c_double), ("kVal", C.c_double), ("rVal", C.c_double),
                ("mean", _PD * 4), ("sd", _PD * 4), ("err", _PD * 4),
                ("F", _PD * 4), ("F1", _PD * 4), ("Gamma", _PD * 4),
                ("D", _PD * 4), ("Dc", _PD * 4), ("H", _PD * 4), ("Hc", _PD * 4),
                ("fourier", C.c_int), ("omegaMax", C.c_int), ("nv", C.c_int), ("tMax", C.c_double)]


# every symbol include/sqt.h declares: (name, restype, argtypes)
ENGINE_API = [
    ("sqt_results_alloc", C.c_int, [C.POINTER(SqtResults), C.c_int, C.c_int]),
    ("sqt_results_alloc_types", C.c_int, [C.POINTER(SqtResults), C.c_int]),
    ("sqt_results_free", None, [C.POINTER(SqtResults)]),
    ("sqt_set_types", C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    ("sqt_multi_set_types", C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    ("sqt_create", C.c_int, [C.POINTER(SqtParams), C.POINTER(C.c_void_p)]),
    ("sqt_destroy", None, [C.c_void_p]),
    ("sqt_last_error", C.c_char_p, [C.c_void_p]),
    ("sqt_set_limit", C.c_int, [C.c_void_p, C.c_int]),
    ("sqt_begin_file", C.c_int, [C.c_void_p]),
    ("sqt_stage_acquire", C.c_void_p, [C.c_void_p, C.c_int]),
    ("sqt_push_frame", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_double, C.c_double]),
    ("sqt_push_frame_xyz", C.c_int, [C.c_void_p] + [C.c_void_p] * 6 + [C.c_int, C.c_longlong, C.c_double, C.c_double]),
    ("sqt_push_frame_xyz_vm", C.c_int, [C.c_void_p] + [C.c_void_p] * 9 + [C.c_int, C.c_longlong, C.c_double, C.c_double]),
    ("sqt_frame_components", C.c_int, [C.c_void_p]),
    ("sqt_push_frame_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_double, C.c_double]),
    ("sqt_push_frame_device_ex", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_double, C.c_double, C.c_void_p]),
    ("sqt_fetch_results", C.c_int, [C.c_void_p, C.POINTER(SqtResults)]),
    ("sqt_finish", C.c_int, [C.c_void_p, C.POINTER(SqtResults)]),
    ("sqt_sync", C.c_int, [C.c_void_p]),
    ("sqt_count_corr_av", C.c_int, [C.c_void_p]),
    ("sqt_buffer_count", C.c_int, [C.c_void_p, C.c_int]),
    ("sqt_number_call_print", C.c_int, [C.c_void_p]),
    ("sqt_get_rho", C.c_int, [C.c_void_p, C.c_void_p]),
    ("sqt_get_kvel", C.c_int, [C.c_void_p, C.c_void_p]),
    ("sqt_get_kmu", C.c_int, [C.c_void_p, C.c_void_p]),
    ("sqt_kernel_launches", C.c_longlong, [C.c_void_p]),
    ("sqt_profile_enable", C.c_int, [C.c_void_p, C.c_int]),
    ("sqt_profile_read", C.c_int, [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_double)]),
    ("sqt_timer_start", C.c_int, [C.c_void_p]),
    ("sqt_timer_stop", C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    ("sqt_kernel_config", C.c_char_p, [C.c_void_p]),
    ("sqt_comm_unique_id", C.c_int, [C.c_void_p]),
    ("sqt_comm_init", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    ("sqt_multi_create", C.c_int, [C.POINTER(SqtParams), C.POINTER(C.c_void_p)]),
    ("sqt_multi_destroy", None, [C.c_void_p]),
    ("sqt_multi_last_error", C.c_char_p, [C.c_void_p]),
    ("sqt_multi_n_gpus", C.c_int, [C.c_void_p]),
    ("sqt_multi_handle", C.c_void_p, [C.c_void_p, C.c_int]),
    ("sqt_multi_set_limit", C.c_int, [C.c_void_p, C.c_int]),
    ("sqt_multi_begin_file", C.c_int, [C.c_void_p]),
    ("sqt_multi_stage_acquire", C.c_void_p, [C.c_void_p, C.c_int]),
    ("sqt_multi_push_frame", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double]),
    ("sqt_multi_fetch_results", C.c_int, [C.c_void_p, C.POINTER(SqtResults)]),
    ("sqt_multi_finish", C.c_int, [C.c_void_p, C.POINTER(SqtResults)]),
    ("sqt_multi_sync", C.c_int, [C.c_void_p]),
    ("sqt_multi_kernel_launches", C.c_longlong, [C.c_void_p]),
    ("sqt_synth_create", C.c_int, [C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_ulonglong, C.POINTER(C.c_void_p)]),
    ("sqt_synth_next", C.c_void_p, [C.c_void_p]),
    ("sqt_synth_box_length", C.c_double, [C.c_void_p]),
    ("sqt_synth_copy_to_host", C.c_int, [C.c_void_p, C.c_void_p]),
    ("sqt_synth_copy_to_device", C.c_int, [C.c_void_p, C.c_void_p]),
    ("sqt_synth_destroy", None, [C.c_void_p]),
    ("sqt_measure_dfma_peak", C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
]

HOST_API = [
    ("sqt_deck_read", C.c_int, [C.c_char_p, C.POINTER(SqtDeck)]),
    ("sqt_deck_print", C.c_int, [C.c_void_p, C.POINTER(SqtDeck)]),
    ("sqt_deck_write", C.c_int, [C.c_char_p, C.POINTER(SqtDeck)]),
    ("sqt_deck_write_template", C.c_int, [C.c_char_p]),
    ("sqt_dump_read_header", C.c_int, [C.c_void_p, C.POINTER(SqtFrameHeader)]),
    ("sqt_dump_read_atoms", C.c_int, [C.c_void_p, C.POINTER(SqtFrameHeader), C.c_void_p, C.c_void_p, C.c_void_p]),
    ("sqt_dump_skip_atoms", C.c_int, [C.c_void_p, C.POINTER(SqtFrameHeader)]),
    ("sqt_reader_open", C.c_void_p, [C.c_char_p, C.c_int]),
    ("sqt_reader_close", None, [C.c_void_p]),
    ("sqt_reader_is_raw", C.c_int, [C.c_void_p]),
    ("sqt_raw_append_frame", C.c_int, [C.c_char_p, C.c_int, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("sqt_reader_rewind", None, [C.c_void_p]),
    ("sqt_reader_next_header", C.c_int, [C.c_void_p, C.POINTER(SqtFrameHeader)]),
    ("sqt_reader_read_atoms", C.c_int, [C.c_void_p, C.POINTER(SqtFrameHeader), C.c_void_p, C.c_void_p, C.c_void_p]),
    ("sqt_reader_skip_atoms", C.c_int, [C.c_void_p, C.POINTER(SqtFrameHeader)]),
    ("sqt_reader_read_atoms_v", C.c_int, [C.c_void_p, C.POINTER(SqtFrameHeader)] + [C.c_void_p] * 6),
    ("sqt_reader_read_atoms_vm", C.c_int, [C.c_void_p, C.POINTER(SqtFrameHeader)] + [C.c_void_p] * 9),
    ("sqt_host_factor_dr", None, [C.POINTER(SqtResults), C.c_double, C.c_void_p]),
    ("sqt_host_normalise", None, [C.POINTER(SqtResults), C.POINTER(SqtDeck), C.c_int, C.c_int, C.c_int]),
    ("sqt_host_print_process", C.c_int, [C.c_char_p, C.POINTER(SqtResults), C.POINTER(SqtDeck), C.c_int, C.c_int, C.c_int, C.c_char_p]),
    ("sqt_host_normalise_v", None, [C.POINTER(SqtResults), C.POINTER(SqtDeck), C.c_int, C.c_int, C.c_int, C.c_int]),
    ("sqt_host_print_process_v", C.c_int, [C.c_char_p, C.POINTER(SqtResults), C.POINTER(SqtDeck), C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p]),
    ("sqt_host_normalise_x", None, [C.POINTER(SqtResults), C.POINTER(SqtDeck), C.POINTER(SqtFlags)]),
    ("sqt_host_print_process_x", C.c_int, [C.c_char_p, C.POINTER(SqtResults), C.POINTER(SqtDeck), C.POINTER(SqtFlags), C.c_char_p]),
    ("sqt_host_normalise_types", None, [C.POINTER(SqtResults), C.POINTER(SqtDeck)]),
    ("sqt_host_print_types", C.c_int, [C.c_char_p, C.POINTER(SqtResults), C.POINTER(SqtDeck), C.POINTER(SqtFlags), C.c_char_p]),
    ("sqt_reader_peek_types", C.c_int, [C.c_void_p, C.POINTER(SqtFrameHeader), C.c_void_p]),
    ("sqt_reader_indexed_frames", C.c_long, [C.c_void_p]),
    ("sqt_prefetch_start", C.c_void_p, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_long]),
    ("sqt_prefetch_next", C.c_int, [C.c_void_p, C.POINTER(SqtFrameHeader), C.POINTER(C.c_void_p)]),
    ("sqt_prefetch_stop", None, [C.c_void_p]),
    ("sqt_analysis_load", C.c_int, [C.POINTER(SqtAnalysis), C.c_char_p, C.c_char_p, C.c_int]),
    ("sqt_analysis_from_results", C.c_int, [C.POINTER(SqtAnalysis), C.POINTER(SqtResults), C.POINTER(SqtDeck), C.c_int]),
    ("sqt_analysis_time", C.c_int, [C.POINTER(SqtAnalysis), C.c_int, C.c_double, C.c_char_p]),
    ("sqt_analysis_fourier", C.c_int, [C.POINTER(SqtAnalysis), C.c_int]),
    ("sqt_analysis_print", C.c_int, [C.POINTER(SqtAnalysis), C.c_char_p]),
    ("sqt_analysis_free", None, [C.POINTER(SqtAnalysis)]),
    ("sqt_analysis_text_load", C.c_int, [C.POINTER(SqtAnalysis), C.c_char_p, C.c_char_p, C.c_int]),
    ("sqt_analysis_text_time", C.c_int, [C.POINTER(SqtAnalysis), C.c_int, C.c_char_p]),
    ("sqt_analysis_text_fourier", C.c_int, [C.POINTER(SqtAnalysis), C.c_int]),
    ("sqt_analysis_text_print", C.c_int, [C.POINTER(SqtAnalysis), C.c_char_p]),
]


def _bind(lib, table):
    for name, res, args in table:
        fn = getattr(lib, name)  # AttributeError here = the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    return lib


_engine = None
_host = None


def engine_lib():
    """The CUDA engine.  Raises if the extension has not been built: there is no CPU fallback."""
    global _engine
    if _engine is None:
        if not os.path.exists(ENGINE_SO):
            raise RuntimeError(
                "lmp2sqt_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C lmp2sqt_b200/csrc`).  The product path is the CUDA path; there is no CPU fallback." % ENGINE_SO)
        _engine = _bind(C.CDLL(ENGINE_SO, mode=C.RTLD_GLOBAL), ENGINE_API)
    return _engine


def host_lib():
    global _host
    if _host is None:
        if not os.path.exists(HOST_SO):
            raise RuntimeError("lmp2sqt_b200: %s is missing -- run `make -C lmp2sqt_b200/csrc`" % HOST_SO)
        _host = _bind(C.CDLL(HOST_SO), HOST_API)
    return _host
