"""ctypes binding of the C ABI declared in include/mnv.h.

The prototypes are parsed from the header itself, so the binding cannot drift from the ABI and a
missing export fails at load time.  There is NO fallback: if the CUDA library is absent or a
symbol is missing this module raises -- the product path never routes through CPU code.
"""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HEADER = os.path.join(ROOT, "include", "mnv.h")
LIB_PATH = os.path.join(HERE, "lib", "libmnv_b200.so")

_CTYPES = {
    "int": C.c_int, "float": C.c_float, "size_t": C.c_size_t, "unsigned int": C.c_uint,
    "uint64_t": C.c_uint64, "mnv_stream_t": C.c_void_p, "void": None,
}


def _ctype(t):
    t = " ".join(t.replace("const", " ").split())
    if t.endswith("*"):
        return C.c_char_p if t.startswith("char") else C.c_void_p
    return _CTYPES[t]


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes], [argnames])} for every function the header declares."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(mnv_\w+)\s*\(([^;{}]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef"):
            continue
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?[\s\*])(\w+)$", a)
                argtypes.append(_ctype(mm.group(1).strip()))
                argnames.append(mm.group(2))
        protos[name] = (_ctype(ret), argtypes, argnames)
    return protos


class MnvError(RuntimeError):
    pass


_lib = None
_tuning = None
PROTOS = parse_header()
TUNING_LIB_PATH = os.path.join(HERE, "lib", "libmnv_b200_tuning.so")


def load():
    """Load libmnv_b200.so and type every entry point.  Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MnvError("CUDA kernel library not built: %s (run `python -m minerva_b200.build`); "
                       "there is no CPU fallback" % LIB_PATH)
    try:
        import torch  # noqa: F401  (loads the CUDA runtime the process will share)
    except Exception:
        pass
    lib = C.CDLL(LIB_PATH)
    for name, (res, argtypes, _) in PROTOS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise MnvError("libmnv_b200.so does not export %s declared in include/mnv.h" % name)
        fn.restype = res
        fn.argtypes = argtypes
    _lib = lib
    return lib


def load_tuning():
    """The TUNING build (include/mnv_debug.h): same ABI plus mnv_debug_set_option.  A separate handle for tools/ and
    the alternate-operand-path tests; nothing in the product (minerva_b200.owl, bench.py's timed path) loads it."""
    global _tuning
    if _tuning is not None:
        return _tuning
    if not os.path.exists(TUNING_LIB_PATH):
        raise MnvError("tuning build not found: %s (python -m minerva_b200.build)" % TUNING_LIB_PATH)
    try:
        import torch  # noqa: F401
    except Exception:
        pass
    lib = C.CDLL(TUNING_LIB_PATH)
    for name, (res, argtypes, _) in PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = argtypes
    lib.mnv_debug_set_option.restype = C.c_int
    lib.mnv_debug_set_option.argtypes = [C.c_char_p, C.c_int]
    _tuning = lib
    return lib


def use_tuning():
    """tools/ only: make the tuning build THE library of this process (load() returns it from now on)."""
    global _lib
    _lib = load_tuning()
    return _lib


_ERRS = {-1: "MNV_EINVAL", -2: "MNV_EUNSUPPORTED", -3: "MNV_EWORKSPACE"}


def check(rc, what):
    """The reference's flat functions return void and CHECK-fail; the host side keeps that."""
    if rc != 0:
        raise MnvError("%s failed: %s" % (what, _ERRS.get(rc, "cudaError %d" % rc)))


def call(name, *args, lib=None):
    fn = getattr(lib or load(), name)
    check(fn(*args), name)
