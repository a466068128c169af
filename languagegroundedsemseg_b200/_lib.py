"""ctypes binding of the C-ABI library (include/lgs_b200.h).  The product path has no CPU fallback:
if the CUDA library is missing or a call fails this module raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "liblgs_b200.so")

OK, E_INVALID, E_CUDA, E_RANGE, E_HASH_FULL, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5
F32, BF16 = 0, 1
ALGO_SIMT, ALGO_TC, ALGO_TC3, ALGO_BX3 = 0, 1, 2, 3
W_KCN, W_KNC, W_KNC_SPLIT, W_BX3 = 0, 1, 2, 3

_p, _i64, _i32, _f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float

# name -> (restype, argtypes); mirrors include/lgs_b200.h one to one
SIGNATURES = {
    "lgs_version": (C.c_int, []),
    "lgs_last_error": (C.c_char_p, []),
    "lgs_launch_count": (C.c_uint64, []),
    "lgs_trace_begin": (C.c_int, []),
    "lgs_trace_end": (_i64, [C.c_char_p, _i64]),
    "lgs_has_tc": (C.c_int, []),
    "lgs_tune": (C.c_int, [C.c_char_p, _i32]),
    "lgs_coord_limit": (_i32, []),
    "lgs_hash_capacity": (_i64, [_i64]),
    "lgs_coordmap_scratch_elems": (_i64, [_i64]),
    "lgs_coordmap_build": (C.c_int, [_p, _i64, _i32, _p, _p, _i64, _p, _p, _p, _p, _p, C.POINTER(_i64), _p]),
    "lgs_kmap_build": (C.c_int, [_p, _i64, _p, _p, _i64, _i32, _i32, _i32, _p, _p, _p]),
    "lgs_kmap_transpose": (C.c_int, [_p, _i32, _i64, _i64, _p, _p]),
    "lgs_conv_tc_supported": (C.c_int, [_i32, _i32, _i32]),
    "lgs_weight_prep": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p, _i32, _p]),
    "lgs_weight_prep_batch": (C.c_int, [_p, _i32, _i64, _i32, _i32, _p]),
    "lgs_conv_fwd": (C.c_int, [_p, _i64, _i32, _p, _i32, _i32, _i32, _p, _i64, _i32, _p, _p, _i32, _i32, _p]),
    "lgs_conv_fwd2": (C.c_int, [_p, _i32, _p, _i32, _i64, _p, _i32, _i32, _p, _i64, _i32, _p, _p, _p, _p]),
    "lgs_conv_fwd3": (C.c_int, [_p, _i32, _p, _i32, _i64, _p, _i32, _i32, _p, _p, _i64, _i32, _p, _p, _p, _p]),
    "lgs_conv_fwd4": (C.c_int, [_p, _i32, _p, _i32, _i64, _p, _i32, _i32, _p, _p, _i64, _i32, _p, _p, _p, _p, _p]),
    "lgs_nbplan_supported": (C.c_int, [_i64, _i32]),
    "lgs_nbplan_bytes": (_i64, [_i64, _i32]),
    "lgs_nbplan_scratch_bytes": (_i64, [_i64]),
    "lgs_nbplan_build": (C.c_int, [_p, _i64, _p, _i32, _i32, _p, _p, _p, _p]),
    "lgs_nbplan_geometry": (C.c_int, [_i64, _i32, C.POINTER(_i64)]),
    "lgs_weight_bx3_elems": (_i64, [_i32, _i32, _i32]),
    "lgs_conv_wgrad": (C.c_int, [_p, _i64, _i32, _p, _i64, _i32, _p, _i32, _p, _i32, _i32, _p]),
    "lgs_bn_fwd": (C.c_int, [_p, _p, _i64, _i32, _p, _p, _f32, _f32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "lgs_bn_fwd2": (C.c_int, [_p, _p, _i64, _i32, _p, _p, _f32, _f32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _i32, _p]),
    "lgs_copy2d": (C.c_int, [_p, _i64, _p, _i64, _i64, _i32, _p]),
    "lgs_add": (C.c_int, [_p, _p, _p, _i64, _p]),
    "lgs_colsum": (C.c_int, [_p, _i64, _i32, _p, _p]),
    "lgs_program_create": (C.c_int, [_p, _i32, _p, _i32, _i32, _i32, C.POINTER(_p)]),
    "lgs_program_destroy": (None, [_p]),
    "lgs_program_arena_bytes": (_i64, [_p, _p]),
    "lgs_program_reset": (None, [_p]),
    "lgs_program_run": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _i64, _p, _p, _p]),
    "lgs_program_run2": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _i64, _p, _p, _p, _i32]),
    "lgs_bn_bwd": (C.c_int, [_p, _p, _p, _i64, _i32, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "lgs_seg_ce_supported": (C.c_int, [_i32]),
    "lgs_seg_ce": (C.c_int, [_p, _i64, _i32, _p, _i64, _p, _p, _p, _p]),
    "lgs_clip_ce": (C.c_int, [_p, _i64, _i32, _p, _i32, _p, _i64, _p, _p, _p, _p, _p]),
    "lgs_clip_ce_tc_supported": (C.c_int, [_i32, _i32]),
    "lgs_clip_ce_tc_ws_elems": (_i64, [_i32, _i32]),
    "lgs_clip_ce_tc": (C.c_int, [_p, _i64, _i32, _p, _i32, _p, _i64, _p, _p, _p, _p, _p, _p]),
    "lgs_clip_hinge": (C.c_int, [_p, _i64, _i32, _p, _i32, _p, _p, _i32, _i64, _f32, _f32, _f32, _p, _p, _p, _p]),
    "lgs_voxelize_affine": (C.c_int, [_p, _i64, C.POINTER(C.c_double), _i32, _p, _p]),
}

_lib = None
_fast = None          # the optional native binding (csrc/_lgs_fast*.so), used when LGS_FAST_BIND=1


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"lgs_b200 error {code}: {msg}")
        self.code = code


class _Entries:
    """namespace of the C-ABI entry points: ctypes functions, overridden by the native binding's where it is enabled"""


def _load_fast():
    import glob
    import importlib.util
    hits = glob.glob(os.path.join(_HERE, "csrc", "_lgs_fast*.so"))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location("_lgs_fast", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    """Load liblgs_b200.so (built in-tree by csrc/build.py).  Raises if it is not there.
    Binding: the int-returning entry points go through csrc/_lgs_fast*.so when it is built — a generated CPython
    extension that calls the same C functions without libffi (1.2 us instead of 6.5 us per call with 15 arguments; a
    training step makes ~480 calls) — and through ctypes otherwise (LGS_FAST_BIND=0 forces ctypes, =1 requires the
    extension).  Both bindings pass the same argument values to the same library: tests/test_host_logic.py and
    tests/test_abi.py compare their recorded calls."""
    global _lib, _fast
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m languagegroundedsemseg_b200.csrc.build` "
                "(or __graft_entry__.build()).  There is no CPU fallback for the engine.")
        lib = C.CDLL(LIB_PATH)
        ns = _Entries()
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
            setattr(ns, name, fn)
        ns._cdll = lib
        want = os.environ.get("LGS_FAST_BIND", "auto")
        if want != "0":
            try:
                _fast = _load_fast()
            except Exception:                      # e.g. built for another interpreter: ctypes serves every entry
                _fast = None
            if _fast is None and want == "1":
                raise RuntimeError("LGS_FAST_BIND=1 but csrc/_lgs_fast*.so is not built (python -m ...csrc.build)")
            if _fast is not None:
                for name in SIGNATURES:
                    if hasattr(_fast, name):
                        setattr(ns, name, getattr(_fast, name))
        # LGS_TUNE="key=value,key=value": lgs_tune knobs for experiments (e.g. pdl=0, nb_off=1); unknown keys raise
        for kv in filter(None, os.environ.get("LGS_TUNE", "").split(",")):
            key, _, val = kv.partition("=")
            if lib.lgs_tune(key.strip().encode(), int(val)) != 0:
                raise RuntimeError(f"LGS_TUNE: unknown knob {key!r}")
        _lib = ns
    return _lib


def binding():
    load()
    return "native" if _fast is not None else "ctypes"


def check(rc):
    if rc != OK:
        raise EngineError(rc, load().lgs_last_error().decode(errors="replace"))


def ptr(t):
    """Raw device pointer of a (contiguous) tensor, or NULL."""
    if _fast is not None:
        return 0 if t is None else t.data_ptr()
    return None if t is None else C.c_void_p(t.data_ptr())


def launch_count():
    return int(load().lgs_launch_count())


class trace:
    """`with _lib.trace() as t: ...; t.lines` — record the compute entry points' calls instead of executing them
    (include/lgs_b200.h: lgs_trace_begin / lgs_trace_end).  Needs no GPU."""

    def __enter__(self):
        load().lgs_trace_begin()
        self.lines = []
        return self

    def __exit__(self, *a):
        lib = load()
        need = lib.lgs_trace_end(None, 0)
        buf = C.create_string_buffer(int(need))
        # recording is already off; the text stays until the next lgs_trace_begin
        lib.lgs_trace_end(buf, need)
        self.lines = buf.value.decode().splitlines()
        return False
