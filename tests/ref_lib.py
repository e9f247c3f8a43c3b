"""ctypes binding of oracle/_ref/libaby3ref.so: the REFERENCE's own sh3 sources (Sh3Runtime, Sh3Encryptor,
Sh3Evaluator, SharedOT, Sh3BinaryEvaluator, Sh3Piecewise, CircuitLibrary) compiled unmodified from
/root/reference against the stand-in third-party headers of oracle/shim (oracle/Makefile target `_ref`,
driver oracle/ref_driver.cpp).  Test infrastructure: used to pin oracle/oracle.cpp, never by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "oracle", "_ref", "libaby3ref.so")
REFERENCE = os.environ.get("ABY3_REFERENCE", "/root/reference")

_p, _u64, _int = C.c_void_p, C.c_uint64, C.c_int


def available():
    """Build when the reference tree is present (this container); otherwise use the prebuilt library
    that travelled with the snapshot (GPU box); otherwise report unavailable."""
    if os.path.isdir(os.path.join(REFERENCE, "aby3", "sh3")):
        r = subprocess.run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "_ref", "REF=" + REFERENCE],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("building oracle/_ref failed:\n" + r.stdout[-4000:])
    return os.path.exists(PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libaby3ref.so is missing and %s is absent" % REFERENCE)
        l = C.CDLL(PATH)
        l.ref_last_error.restype = C.c_char_p
        l.ref_session_new.restype = _p
        l.ref_session_new.argtypes = [_p, _p]
        l.ref_session_free.argtypes = [_p]
        l.ref_session_set_disable_randomization.argtypes = [_p, _int]
        l.ref_share_int.argtypes = [_p, _int, _p, _p, _u64, _u64]
        l.ref_share_bin.argtypes = [_p, _int, _p, _p, _u64, _u64]
        l.ref_reveal_all.argtypes = [_p, _p, _u64, _u64, _int, _p]
        l.ref_mul.argtypes = [_p, _p, _p, _p, _u64, _u64]
        l.ref_mul_trunc.argtypes = [_p, _p, _p, _p, _u64, _u64, _u64]
        l.ref_trunc_tuple.argtypes = [_p, _int, _u64, _u64, _u64, _p, _p, _p]
        l.ref_mul_bit.argtypes = [_p, _p, _p, _p, _u64]
        l.ref_mul_bit_pub.argtypes = [_p, C.c_int64, _p, _p, _u64]
        l.ref_bin_eval.argtypes = [_p, _p, _u64, _p, _p]
        l.ref_piecewise.argtypes = [_p, _p, _u64, _p, _int, _p, _p, _p, _p, _u64, _p]
        l.ref_piecewise_plain.argtypes = [_p, _u64, _p, _int, _p, _p, _p, _p, _u64, _p]
        l.ref_conv_init.argtypes = [_p]
        l.ref_conv_a2b.argtypes = [_p, _p, _u64, _u64, _p]
        l.ref_conv_bit_injection.argtypes = [_p, _p, _u64, _u64, _p]
        l.ref_conv_packed_roundtrip.argtypes = [_p, _u64, _u64, _p, _u64, _p]
        l.ref_share_reveal_packed.argtypes = [_p, _int, _p, _u64, _u64, _p, _p]
        l.ref_time_mul_trunc.restype = C.c_double
        l.ref_time_mul_trunc.argtypes = [_p, _u64, _u64, _u64, _u64, _int]
        l.ref_time_logistic.restype = C.c_double
        l.ref_time_logistic.argtypes = [_p, _u64, _u64, _u64, _int]
        l.ref_main_linear.argtypes = [_int, _p]
        l.ref_sgd_linear.restype = C.c_double
        l.ref_sgd_linear.argtypes = [_p, _p, _u64, _u64, _u64, _u64, C.c_double, _p]
        l.ref_basic_bool.argtypes = [_p, _int, _p, _p, _u64, _p, _p]
        l.ref_basic_cipher_gt.argtypes = [_p, _p, _p, _u64, _p, _p]
        l.ref_basic_max_min_split.argtypes = [_p, _p, _p, _u64, _p, _p, _p]
        l.ref_basic_odd_even_merge.argtypes = [_p, _p, _u64, _p, _u64, _p, _p]
        l.ref_role_test.argtypes = [C.c_char_p, _p, _p]
        _lib = l
    return _lib


def ptr(a):
    return a.ctypes.data_as(_p)


def _chk(rc):
    if rc != 0:
        raise RuntimeError("reference: " + lib().ref_last_error().decode())


def _coefs(thresholds, coefficients):
    th = np.asarray(thresholds, dtype=np.float64)
    counts = np.asarray([len(c) for c in coefficients], dtype=np.int32)
    flat = [v for c in coefficients for v in c]
    is_int = np.asarray([isinstance(v, (int, np.integer)) for v in flat] or [0], dtype=np.int32)
    ints = np.asarray([int(v) if isinstance(v, (int, np.integer)) else 0 for v in flat] or [0], dtype=np.int64)
    dbl = np.asarray([float(v) for v in flat] or [0.0], dtype=np.float64)
    return th, counts, is_int, ints, dbl


class Session:
    """Three reference parties (threads) with the seeds of aby3_tests/Sh3EvaluatorTests.cpp:41-47 by default."""

    def __init__(self, enc_seeds, eval_seeds):
        self._e = C.create_string_buffer(enc_seeds, 96)
        self._v = C.create_string_buffer(eval_seeds, 96)
        self.h = lib().ref_session_new(C.cast(self._e, _p), C.cast(self._v, _p))
        if not self.h:
            raise RuntimeError("reference: " + lib().ref_last_error().decode())

    def close(self):
        if self.h:
            lib().ref_session_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def disable_randomization(self, on=True):
        lib().ref_session_set_disable_randomization(self.h, int(on))

    def share_int(self, owner, plain):
        plain = np.ascontiguousarray(plain, dtype=np.int64)
        sh = np.empty((3, 2) + plain.shape, dtype=np.int64)
        _chk(lib().ref_share_int(self.h, owner, ptr(plain), ptr(sh), plain.shape[0], plain.shape[1]))
        return sh

    def share_bin(self, owner, plain):
        plain = np.ascontiguousarray(plain, dtype=np.int64)
        sh = np.empty((3, 2) + plain.shape, dtype=np.int64)
        _chk(lib().ref_share_bin(self.h, owner, ptr(plain), ptr(sh), plain.shape[0], plain.shape[1]))
        return sh

    def reveal_all(self, shares, binary=False):
        shares = np.ascontiguousarray(shares, dtype=np.int64)
        r, c = shares.shape[2], shares.shape[3]
        out = np.empty((3, r, c), dtype=np.int64)
        _chk(lib().ref_reveal_all(self.h, ptr(shares), r, c, int(binary), ptr(out)))
        return out

    def mul(self, A, B):
        A, B = np.ascontiguousarray(A), np.ascontiguousarray(B)
        Cc = np.empty_like(A)
        _chk(lib().ref_mul(self.h, ptr(A), ptr(B), ptr(Cc), A.shape[2], A.shape[3]))
        return Cc

    def mul_trunc(self, A, B, shift):
        A, B = np.ascontiguousarray(A), np.ascontiguousarray(B)
        Cc = np.empty_like(A)
        _chk(lib().ref_mul_trunc(self.h, ptr(A), ptr(B), ptr(Cc), A.shape[2], A.shape[3], shift))
        return Cc

    def trunc_tuple(self, party, rows, cols, d):
        n = rows * cols
        R, T0, T1 = (np.empty(n, dtype=np.int64) for _ in range(3))
        _chk(lib().ref_trunc_tuple(self.h, party, rows, cols, d, ptr(R), ptr(T0), ptr(T1)))
        return R, T0, T1

    def mul_bit(self, A, B):
        A, B = np.ascontiguousarray(A), np.ascontiguousarray(B)
        n = A[0, 0].size
        Cc = np.empty((3, 2, n, 1), dtype=np.int64)
        _chk(lib().ref_mul_bit(self.h, ptr(A), ptr(B), ptr(Cc), n))
        return Cc

    def mul_bit_pub(self, a, B):
        B = np.ascontiguousarray(B)
        n = B[0, 0].size
        Cc = np.empty((3, 2, n, 1), dtype=np.int64)
        _chk(lib().ref_mul_bit_pub(self.h, int(a), ptr(B), ptr(Cc), n))
        return Cc

    def bin_eval(self, cir, width, inputs):
        import oracle_lib as o
        c = o.Circuit()
        keep = []

        def arr(a):
            keep.append(a)
            return a.ctypes.data_as(_p)

        c.wire_count, c.gate_count = cir["wire_count"], len(cir["gates"]) // 4
        c.gates, c.level_count, c.level_gates = arr(cir["gates"]), len(cir["level_gates"]), arr(cir["level_gates"])
        c.num_inputs, c.input_first, c.input_bits = len(cir["input_bits"]), arr(cir["input_first"]), arr(cir["input_bits"])
        c.num_outputs, c.output_off, c.output_bits = len(cir["output_bits"]), arr(cir["output_off"]), arr(cir["output_bits"])
        c.output_wires, c.output_invert = arr(cir["output_wires"]), arr(cir["output_invert"])
        ins = [np.ascontiguousarray(x, dtype=np.int64) for x in inputs]
        outs = [np.zeros((3, 2, width, (int(b) + 63) // 64), dtype=np.int64) for b in cir["output_bits"]]
        in_ptrs = (_p * len(ins))(*[x.ctypes.data_as(_p) for x in ins])
        out_ptrs = (_p * len(outs))(*[x.ctypes.data_as(_p) for x in outs])
        _chk(lib().ref_bin_eval(self.h, C.byref(c), width, in_ptrs, out_ptrs))
        return outs

    def piecewise(self, X, thresholds, coefficients, D):
        X = np.ascontiguousarray(X, dtype=np.int64)
        n = X[0, 0].size
        th, counts, is_int, ints, dbl = _coefs(thresholds, coefficients)
        Y = np.empty((3, 2, n, 1), dtype=np.int64)
        _chk(lib().ref_piecewise(self.h, ptr(X), n, ptr(th), len(th), ptr(counts), ptr(is_int), ptr(ints), ptr(dbl), D, ptr(Y)))
        return Y

    def conv_init(self):
        _chk(lib().ref_conv_init(self.h))

    def conv_a2b(self, X):
        X = np.ascontiguousarray(X, dtype=np.int64)
        Y = np.empty_like(X)
        _chk(lib().ref_conv_a2b(self.h, ptr(X), X.shape[2], X.shape[3], ptr(Y)))
        return Y

    def conv_bit_injection(self, B, bits):
        B = np.ascontiguousarray(B, dtype=np.int64)
        rows = B.shape[2]
        Y = np.empty((3, 2, rows, bits), dtype=np.int64)
        _chk(lib().ref_conv_bit_injection(self.h, ptr(B), rows, bits, ptr(Y)))
        return Y

    def share_reveal_packed(self, owner, plain):
        """localPackedBinary / remotePackedBinary then revealAll(sPackedBin): (shares [3][2][bits][simd], revealed [3][rows][cols])"""
        plain = np.ascontiguousarray(plain, dtype=np.int64)
        rows, cols = plain.shape
        sh = np.zeros((3, 2, 64 * cols, (rows + 63) // 64), dtype=np.int64)
        rev = np.zeros((3, rows, cols), dtype=np.int64)
        _chk(lib().ref_share_reveal_packed(self.h, owner, ptr(plain), rows, cols, ptr(sh), ptr(rev)))
        return sh, rev

    def time_mul_trunc(self, M, K, N, shift, reps=1):
        """seconds per asyncMul(A (M x K), B (K x N), C, shift).get() over three party threads"""
        t = lib().ref_time_mul_trunc(self.h, M, K, N, shift, reps)
        if t < 0:
            raise RuntimeError("reference: " + lib().ref_last_error().decode())
        return t


BOOL_OPS = {"lt": 0, "eq": 1, "and": 2, "or": 3, "add": 4, "max": 5, "min": 6}


def _basic(self, fn, *args):
    secs = C.c_double(0)
    _chk(fn(self.h, *args, C.byref(secs)))
    return secs.value


def _session_basic_methods():
    def basic_bool(self, op, A, B):
        """aby3-Basic bool_cipher_<op> (BoolBasic.cpp) on binary sharings [3][2][n][1] -> (out, seconds)"""
        A, B = np.ascontiguousarray(A, dtype=np.int64), np.ascontiguousarray(B, dtype=np.int64)
        n = A.shape[2]
        out = np.zeros((3, 2, n, 1), dtype=np.int64)
        t = _basic(self, lib().ref_basic_bool, BOOL_OPS[op], ptr(A), ptr(B), n, ptr(out))
        return out, t

    def cipher_gt(self, A, B):
        A, B = np.ascontiguousarray(A, dtype=np.int64), np.ascontiguousarray(B, dtype=np.int64)
        n = A.shape[2]
        out = np.zeros((3, 2, n, 1), dtype=np.int64)
        t = _basic(self, lib().ref_basic_cipher_gt, ptr(A), ptr(B), n, ptr(out))
        return out, t

    def max_min_split(self, A, B):
        A, B = np.ascontiguousarray(A, dtype=np.int64), np.ascontiguousarray(B, dtype=np.int64)
        n = A.shape[2]
        mx, mn = np.zeros((3, 2, n, 1), dtype=np.int64), np.zeros((3, 2, n, 1), dtype=np.int64)
        t = _basic(self, lib().ref_basic_max_min_split, ptr(A), ptr(B), n, ptr(mx), ptr(mn))
        return mx, mn, t

    def odd_even_merge(self, A, B):
        A, B = np.ascontiguousarray(A, dtype=np.int64), np.ascontiguousarray(B, dtype=np.int64)
        n1, n2 = A.shape[2], B.shape[2]
        out = np.zeros((3, 2, n1 + n2, 1), dtype=np.int64)
        t = _basic(self, lib().ref_basic_odd_even_merge, ptr(A), n1, ptr(B), n2, ptr(out))
        return out, t

    for f in (basic_bool, cipher_gt, max_min_split, odd_even_merge):
        setattr(Session, f.__name__, f)


_session_basic_methods()


def sgd_linear(x, y, batch, iters, lr=2.0 ** -10, want_shares=True):
    """The reference's aby3-ML linear regression (aby3ML engine + Regression.h SGD_Linear, D16) on x (N x F doubles),
    y (N) -> (seconds for `iters` iterations, w shares [3][2][F][1] or None)"""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
    N, F = x.shape
    w = np.zeros((3, 2, F, 1), dtype=np.int64) if want_shares else None
    t = lib().ref_sgd_linear(ptr(x), ptr(y), N, F, batch, iters, float(lr), ptr(w) if want_shares else None)
    if t < 0:
        raise RuntimeError("reference: " + lib().ref_last_error().decode())
    return t, w


def main_linear(*args):
    """aby3-ML/main-linear.cpp's linear_main_3pc_sh(CLP&) with argv-style arguments, e.g. main_linear("-N", 1000, "-D", 100)"""
    argv = [b"main-linear"] + [str(a).encode() for a in args]
    arr = (C.c_char_p * len(argv))(*argv)
    _chk(lib().ref_main_linear(len(argv), arr))


def _time_logistic(self, rows, features, D=16, reps=1):
    """seconds per pass of y = logisticFunc(X * W) on the reference's own code (X rows x features)"""
    t = lib().ref_time_logistic(self.h, rows, features, D, reps)
    if t < 0:
        raise RuntimeError("reference: " + lib().ref_last_error().decode())
    return t


Session.time_logistic = _time_logistic


def piecewise_plain(x, thresholds, coefficients, D):
    x = np.ascontiguousarray(x, dtype=np.int64).reshape(-1)
    th, counts, is_int, ints, dbl = _coefs(thresholds, coefficients)
    y = np.empty_like(x)
    _chk(lib().ref_piecewise_plain(ptr(x), x.size, ptr(th), len(th), ptr(counts), ptr(is_int), ptr(ints), ptr(dbl), D, ptr(y)))
    return y


def conv_packed_roundtrip(planes, bits):
    """Sh3Converter::toPackedBin + toBinaryMatrix(sPackedBin) of the reference on one pair of share planes
    [2][rows][words] -> (packed [2][bits][simd], unpacked [2][rows][words])"""
    planes = np.ascontiguousarray(planes, dtype=np.int64)
    rows = planes.shape[1]
    simd = (rows + 63) // 64
    packed = np.zeros((2, bits, simd), dtype=np.int64)
    out = np.zeros_like(planes)
    _chk(lib().ref_conv_packed_roundtrip(ptr(planes), rows, bits, ptr(packed), simd, ptr(out)))
    return packed, out


def role_test(name):
    """One of the fork's own role tests (aby3_tests/Test.cpp, BoolTest.cpp, SortTest.cpp), compiled unmodified, on three
    threads with "-role i" -> (number of check_result SUCCESS lines, number of ERROR lines)"""
    ok, bad = C.c_int(0), C.c_int(0)
    _chk(lib().ref_role_test(name.encode(), C.byref(ok), C.byref(bad)))
    return ok.value, bad.value
