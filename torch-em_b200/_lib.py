"""ctypes binding of ``libb200em.so`` -- the C ABI declared in ``include/b200em.h``.

The prototypes are parsed from the header itself, so the Python side can never drift from the C side, and
``declared_symbols()`` is what the CPU test-suite checks the library exports.  There is deliberately NO fallback:
if the library is missing or a call fails, a ``RuntimeError`` is raised (OOM messages keep the words
"out of memory" so torch-em's ``util/memory.py:16-21`` OOM detection keeps working).
"""
import ctypes
import os
import re
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "b200em.h")
LIB_PATH = os.path.join(HERE, "libb200em.so")

F32, BF16 = 0, 1
ACT = {None: 0, "none": 0, "Sigmoid": 1, "ReLU": 2, "Tanh": 3}

_CT = {
    "int": ctypes.c_int, "int64_t": ctypes.c_int64, "float": ctypes.c_float, "void": None,
}


def _ctype(decl):
    decl = decl.strip()
    if "*" in decl:
        return ctypes.c_char_p if decl.replace(" ", "").startswith("constchar*") else ctypes.c_void_p
    base = decl.replace("const", "").split()[0]
    return _CT[base]


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes], [argnames])} for every ``b200em_*`` prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(?:^|\n)\s*((?:const\s+)?\w+\s*\*?)\s*(b200em_\w+)\s*\(([^;{]*?)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        args = " ".join(args.split())
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                nm = re.search(r"(\w+)$", a).group(1)
                argtypes.append(_ctype(a[: -len(nm)]))
                argnames.append(nm)
        protos[name] = (_ctype(ret + " "), argtypes, argnames)
    return protos


def declared_symbols():
    return sorted(parse_header())


_lock = threading.Lock()
_lib = None


def load():
    """Load (once) and return the ctypes library with prototypes set.  Raises if the .so is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python torch-em_b200/build.py` "
                "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback for this path.")
        import torch  # noqa: F401  (loads libcudart.so.12 first so both sides share one runtime)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (ret, argtypes, _) in parse_header().items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = ret
            fn.argtypes = argtypes
        if lib.b200em_abi_version() != 1:
            raise RuntimeError("libb200em.so ABI version mismatch with include/b200em.h")
        _lib = lib
    return _lib


def check(status, what):
    if status != 0:
        msg = load().b200em_last_error().decode(errors="replace")
        raise RuntimeError(f"b200em {what} failed (status {status}): {msg}")


def call(name, *args):
    """Call an int-returning entry point and raise on a non-zero status."""
    check(getattr(load(), name)(*args), name)


def launch_count():
    return int(load().b200em_launch_count())


def reset_launch_count():
    load().b200em_reset_launch_count()
