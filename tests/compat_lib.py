"""ctypes binding of compat/_build/libcompat_apps.so: the REFERENCE's application sources (aby3-ML, aby3-Basic) compiled
unmodified from /root/reference against compat/include (forwarding headers onto the B200 facade) -- compat/Makefile,
driver compat/apps_driver.cpp.  The same entry-point shapes as tests/ref_lib.py (the same sources on the CPU)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "compat", "_build", "libcompat_apps.so")
REFERENCE = os.environ.get("ABY3_REFERENCE", "/root/reference")
APP_SOURCES = ["aby3-ML/aby3ML.cpp", "aby3-ML/LinearModelGen.cpp", "aby3-ML/main-linear.cpp", "aby3-ML/Regression.h", "aby3-ML/aby3ML.h",
               "aby3-Basic/BoolBasic.cpp", "aby3-Basic/ArithBasic.cpp", "aby3-Basic/BuildingBlocks.cpp", "aby3-Basic/Sort.cpp",
               "aby3-Basic/Basic.cpp", "aby3-Basic/debug.cpp", "aby3-Basic/Shuffle.cpp", "aby3-Basic/Basics.h",
               "aby3-Basic/BuildingBlocks.h", "aby3_tests/Test.cpp", "aby3_tests/BoolTest.cpp", "aby3_tests/SortTest.cpp",
               "aby3_tests/Test.h"]

# the fork's own role tests (aby3_tests/Test.h; frontend/main.cpp:16-54 dispatches them by flag) -> check_result() lines each
# one writes per party 0 run, as counted from the reference's own CPU build (tests/test_ref_parity.py pins these)
ROLE_TESTS = {"arith_basic_test": 7, "bool_basic_test": 17, "bool_basic_test2": 6, "bool_aggregation_test": 2,
              "get_first_zero_test": 2, "share_conversion_test": 2, "initialization_test": 6, "bc_sort_test": 1,
              "bc_sort_corner_test": 1, "bc_sort_multiple_times": 1, "quick_sort_test": 1, "quick_sort_with_duplicate_elements_test": 1, "odd_even_merge_test": 6,
              "shuffle_test": 3, "correlation_test": 6}

_p, _u64, _int = C.c_void_p, C.c_uint64, C.c_int


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, "aby3-ML"))


def build():
    """(re)build from the reference tree; returns make's output.  Raises when the build fails."""
    from aby3_b200 import build as b
    b.build_all()
    r = subprocess.run(["make", "-j8", "-C", os.path.join(ROOT, "compat"), "REF=" + REFERENCE],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("building compat/_build failed:\n" + r.stdout[-6000:])
    return r.stdout


def available():
    if have_reference():
        build()
    return os.path.exists(PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("compat/_build/libcompat_apps.so is missing and %s is absent" % REFERENCE)
        l = C.CDLL(PATH)
        l.cmp_last_error.restype = C.c_char_p
        l.cmp_session_new.restype = _p
        l.cmp_session_new.argtypes = [_p, _p]
        l.cmp_session_free.argtypes = [_p]
        l.cmp_share_bin.argtypes = [_p, _int, _p, _p, _u64, _u64]
        l.cmp_share_int.argtypes = [_p, _int, _p, _p, _u64, _u64]
        l.cmp_reveal_all.argtypes = [_p, _p, _u64, _u64, _int, _p]
        l.cmp_basic_bool.argtypes = [_p, _int, _p, _p, _u64, _p, _p]
        l.cmp_basic_cipher_gt.argtypes = [_p, _p, _p, _u64, _p, _p]
        l.cmp_basic_max_min_split.argtypes = [_p, _p, _p, _u64, _p, _p, _p]
        l.cmp_basic_odd_even_merge.argtypes = [_p, _p, _u64, _p, _u64, _p, _p]
        l.cmp_basic_cipher_mul.argtypes = [_p, _p, _p, _u64, _p]
        l.cmp_main_linear.argtypes = [_int, _p]
        l.cmp_sgd_linear.restype = C.c_double
        l.cmp_sgd_linear.argtypes = [_p, _p, _u64, _u64, _u64, _u64, C.c_double, _p]
        l.cmp_role_test.argtypes = [C.c_char_p, _p, _p]
        _lib = l
    return _lib


def ptr(a):
    return a.ctypes.data_as(_p)


def _chk(rc):
    if rc != 0:
        raise RuntimeError("compat: " + lib().cmp_last_error().decode())


BOOL_OPS = {"lt": 0, "eq": 1, "and": 2, "or": 3, "add": 4, "max": 5, "min": 6}


class Session:
    """Three parties (threads), each with the reference's Sh3Runtime / Sh3Encryptor / Sh3Evaluator API on the facade."""

    def __init__(self, enc_seeds, eval_seeds):
        self._e = C.create_string_buffer(enc_seeds, 96)
        self._v = C.create_string_buffer(eval_seeds, 96)
        self.h = lib().cmp_session_new(C.cast(self._e, _p), C.cast(self._v, _p))
        if not self.h:
            raise RuntimeError("compat: " + lib().cmp_last_error().decode())

    def close(self):
        if self.h:
            lib().cmp_session_free(self.h)
            self.h = None

    def share_bin(self, owner, plain):
        plain = np.ascontiguousarray(plain, dtype=np.int64)
        sh = np.empty((3, 2) + plain.shape, dtype=np.int64)
        _chk(lib().cmp_share_bin(self.h, owner, ptr(plain), ptr(sh), plain.shape[0], plain.shape[1]))
        return sh

    def share_int(self, owner, plain):
        plain = np.ascontiguousarray(plain, dtype=np.int64)
        sh = np.empty((3, 2) + plain.shape, dtype=np.int64)
        _chk(lib().cmp_share_int(self.h, owner, ptr(plain), ptr(sh), plain.shape[0], plain.shape[1]))
        return sh

    def reveal_all(self, shares, binary=False):
        shares = np.ascontiguousarray(shares, dtype=np.int64)
        r, c = shares.shape[2], shares.shape[3]
        out = np.empty((3, r, c), dtype=np.int64)
        _chk(lib().cmp_reveal_all(self.h, ptr(shares), r, c, int(binary), ptr(out)))
        return out

    def _timed(self, fn, *args):
        secs = C.c_double(0)
        _chk(fn(self.h, *args, C.byref(secs)))
        return secs.value

    def basic_bool(self, op, A, B):
        A, B = np.ascontiguousarray(A, dtype=np.int64), np.ascontiguousarray(B, dtype=np.int64)
        n = A.shape[2]
        out = np.zeros((3, 2, n, 1), dtype=np.int64)
        return out, self._timed(lib().cmp_basic_bool, BOOL_OPS[op], ptr(A), ptr(B), n, ptr(out))

    def cipher_gt(self, A, B):
        A, B = np.ascontiguousarray(A, dtype=np.int64), np.ascontiguousarray(B, dtype=np.int64)
        n = A.shape[2]
        out = np.zeros((3, 2, n, 1), dtype=np.int64)
        return out, self._timed(lib().cmp_basic_cipher_gt, ptr(A), ptr(B), n, ptr(out))

    def max_min_split(self, A, B):
        A, B = np.ascontiguousarray(A, dtype=np.int64), np.ascontiguousarray(B, dtype=np.int64)
        n = A.shape[2]
        mx, mn = np.zeros((3, 2, n, 1), dtype=np.int64), np.zeros((3, 2, n, 1), dtype=np.int64)
        t = self._timed(lib().cmp_basic_max_min_split, ptr(A), ptr(B), n, ptr(mx), ptr(mn))
        return mx, mn, t

    def odd_even_merge(self, A, B):
        A, B = np.ascontiguousarray(A, dtype=np.int64), np.ascontiguousarray(B, dtype=np.int64)
        n1, n2 = A.shape[2], B.shape[2]
        out = np.zeros((3, 2, n1 + n2, 1), dtype=np.int64)
        return out, self._timed(lib().cmp_basic_odd_even_merge, ptr(A), n1, ptr(B), n2, ptr(out))

    def cipher_mul(self, A, B):
        A, B = np.ascontiguousarray(A, dtype=np.int64), np.ascontiguousarray(B, dtype=np.int64)
        n = A.shape[2]
        out = np.zeros((3, 2, n, 1), dtype=np.int64)
        _chk(lib().cmp_basic_cipher_mul(self.h, ptr(A), ptr(B), n, ptr(out)))
        return out


def sgd_linear(x, y, batch, iters, lr=2.0 ** -10):
    """the reference's aby3ML engine + Regression.h SGD_Linear on the facade -> (seconds, w shares [3][2][F][1])"""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
    N, F = x.shape
    w = np.zeros((3, 2, F, 1), dtype=np.int64)
    t = lib().cmp_sgd_linear(ptr(x), ptr(y), N, F, batch, iters, float(lr), ptr(w))
    if t < 0:
        raise RuntimeError("compat: " + lib().cmp_last_error().decode())
    return t, w


def main_linear(*args):
    argv = [b"main-linear"] + [str(a).encode() for a in args]
    arr = (C.c_char_p * len(argv))(*argv)
    _chk(lib().cmp_main_linear(len(argv), arr))


def role_test(name):
    """One of the fork's own role tests (aby3_tests/Test.cpp, BoolTest.cpp, SortTest.cpp), compiled unmodified, on three
    threads with "-role i" -> (number of check_result SUCCESS lines, number of ERROR lines)"""
    ok, bad = C.c_int(0), C.c_int(0)
    _chk(lib().cmp_role_test(name.encode(), C.byref(ok), C.byref(bad)))
    return ok.value, bad.value
