"""ctypes binding of oracle/liboracle.so -- the CPU checker.  Lives under tests/
because only tests, smoke() and bench.py's CPU-baseline legs may load the oracle."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "oracle", "liboracle.so")

_p, _u64, _int = C.c_void_p, C.c_uint64, C.c_int


def _load():
    if not os.path.exists(PATH):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    l = C.CDLL(PATH)
    l.orc_prng_new.restype = _p
    l.orc_prng_new.argtypes = [C.c_char_p, _u64]
    l.orc_prng_free.argtypes = [_p]
    l.orc_prng_get.argtypes = [_p, _p, _u64]
    l.orc_prng_bytes_consumed.restype = _u64
    l.orc_prng_bytes_consumed.argtypes = [_p]
    l.orc_session_new.restype = _p
    l.orc_session_new.argtypes = [_p, _p]
    l.orc_session_free.argtypes = [_p]
    l.orc_session_set_disable_randomization.argtypes = [_p, _int]
    l.orc_session_cursors.argtypes = [_p, _int, _p]
    l.orc_aes128_encrypt.argtypes = [C.c_char_p, C.c_char_p, _p, _int]
    l.orc_aes_ctr_blocks.argtypes = [C.c_char_p, _u64, _u64, _p]
    l.orc_keystream.argtypes = [C.c_char_p, _u64, _u64, _p]
    l.orc_share_int.argtypes = [_p, _int, _p, _p, _u64]
    l.orc_share_bin.argtypes = [_p, _int, _p, _p, _u64]
    l.orc_reveal.argtypes = [_p, _u64, _int, _int, _p]
    l.orc_mul.argtypes = [_p, _p, _p, _p, _u64, _u64, _u64, _int, _int]
    l.orc_mul_trunc.argtypes = [_p, _p, _p, _p, _u64, _u64, _u64, _int, _u64, _int]
    l.orc_trunc_tuple.argtypes = [_p, _int, _u64, _u64, _p, _p, _p]
    l.orc_mul_bit.argtypes = [_p, _p, _p, _p, _u64]
    l.orc_mul_bit_pub.argtypes = [_p, C.c_int64, _p, _p, _u64]
    l.orc_share_op.argtypes = [_p, _p, _p, _u64, _int]
    l.orc_plain_mul.argtypes = [_p, _p, _p, _u64, _u64, _u64, _int, _int]
    l.orc_cross_term.argtypes = [_p, _p, _p, _p, _p, _u64, _u64, _u64, _int, _int]
    l.orc_bit_transpose.argtypes = [_p, _u64, _u64, _u64, _p, _u64]
    l.orc_bin_row_bytes.restype = _u64
    l.orc_bin_row_bytes.argtypes = [_u64]
    l.orc_bin_eval.argtypes = [_p, _p, _u64, _p, _p, _p]
    l.orc_conv_init.argtypes = [_p]
    l.orc_conv_a2b_inputs.argtypes = [_p, _p, _u64, _p, _p]
    l.orc_conv_bit_injection.argtypes = [_p, _p, _u64, _u64, _u64, _p]
    return l


lib = _load()


def ptr(a):
    return a.ctypes.data_as(_p)


def to_block(hi, lo):
    """oc::toBlock(hi, lo) = _mm_set_epi64x(hi, lo): bytes 0..7 = lo LE, 8..15 = hi LE."""
    return int(lo & (2**64 - 1)).to_bytes(8, "little") + int(hi & (2**64 - 1)).to_bytes(8, "little")


def keystream(key, off, n):
    out = np.empty(n, dtype=np.uint8)
    lib.orc_keystream(bytes(key), off, n, ptr(out))
    return out


def stream_u64(key, e0, n):
    return keystream(key, 8 * e0, 8 * n).view(np.uint64)


def default_seeds():
    """aby3_tests/Sh3EvaluatorTests.cpp:41-47: enc seeds toBlock(0,i), eval seeds toBlock(1,i);
    party i is initialised with (prev = i, next = i+1)."""
    enc = b"".join(to_block(0, i) + to_block(0, (i + 1) % 3) for i in range(3))
    ev = b"".join(to_block(1, i) + to_block(1, (i + 1) % 3) for i in range(3))
    return enc, ev


class Session:
    def __init__(self, enc_seeds=None, eval_seeds=None):
        e, v = default_seeds()
        self.enc_seeds = enc_seeds or e
        self.eval_seeds = eval_seeds or v
        self._e = C.create_string_buffer(self.enc_seeds, 96)
        self._v = C.create_string_buffer(self.eval_seeds, 96)
        self.h = lib.orc_session_new(C.cast(self._e, _p), C.cast(self._v, _p))

    def close(self):
        if self.h:
            lib.orc_session_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def seed(self, kind, party, which):
        """kind 'enc'|'eval', which 0=prev 1=next -> 16 bytes"""
        s = self.enc_seeds if kind == "enc" else self.eval_seeds
        o = (party * 2 + which) * 16
        return s[o:o + 16]

    def disable_randomization(self, on=True):
        lib.orc_session_set_disable_randomization(self.h, int(on))

    def cursors(self, party):
        c = np.zeros(6, dtype=np.uint64)
        lib.orc_session_cursors(self.h, party, ptr(c))
        return c

    def share_int(self, owner, plain):
        plain = np.ascontiguousarray(plain, dtype=np.int64)
        sh = np.empty((3, 2) + plain.shape, dtype=np.int64)
        lib.orc_share_int(self.h, owner, ptr(plain), ptr(sh), plain.size)
        return sh

    def share_bin(self, owner, plain):
        plain = np.ascontiguousarray(plain, dtype=np.int64)
        sh = np.empty((3, 2) + plain.shape, dtype=np.int64)
        lib.orc_share_bin(self.h, owner, ptr(plain), ptr(sh), plain.size)
        return sh

    def mul(self, A, B, mode=0, nthreads=1):
        M, K, N = _dims(A, B, mode)
        Cc = np.empty((3, 2, M, N), dtype=np.int64)
        lib.orc_mul(self.h, ptr(A), ptr(B), ptr(Cc), M, K, N, mode, nthreads)
        return Cc

    def mul_trunc(self, A, B, shift, mode=0, nthreads=1):
        M, K, N = _dims(A, B, mode)
        Cc = np.empty((3, 2, M, N), dtype=np.int64)
        lib.orc_mul_trunc(self.h, ptr(A), ptr(B), ptr(Cc), M, K, N, mode, shift, nthreads)
        return Cc

    def mul_bit(self, A, B):
        n = A[0, 0].size
        Cc = np.empty((3, 2, n, 1), dtype=np.int64)
        lib.orc_mul_bit(self.h, ptr(np.ascontiguousarray(A)), ptr(np.ascontiguousarray(B)), ptr(Cc), n)
        return Cc

    def mul_bit_pub(self, a, B):
        n = B[0, 0].size
        Cc = np.empty((3, 2, n, 1), dtype=np.int64)
        lib.orc_mul_bit_pub(self.h, int(a), ptr(np.ascontiguousarray(B)), ptr(Cc), n)
        return Cc

    def trunc_tuple(self, party, n, d):
        R, T0, T1 = (np.empty(n, dtype=np.int64) for _ in range(3))
        lib.orc_trunc_tuple(self.h, party, n, d, ptr(R), ptr(T0), ptr(T1))
        return R, T0, T1

    def share_packed(self, owner, plain):
        """Sh3Encryptor::localPackedBinary / remotePackedBinary (Sh3Encryptor.cpp:342-425): the plaintext rows are
        bit-transposed (one row per bit, one bit per secret) and then shared exactly like localBinMatrix."""
        plain = np.ascontiguousarray(plain, dtype=np.int64)
        rows, cols = plain.shape
        simd = (rows + 63) // 64
        t = bit_transpose(plain.view(np.uint8).reshape(-1), rows, 64 * cols, cols * 8, simd * 8)
        return self.share_bin(owner, t.view(np.int64).reshape(64 * cols, simd))

    def conv_init(self):
        """Sh3Converter::init(rt, eval.mShareGen) on every party"""
        lib.orc_conv_init(self.h)

    def conv_a2b(self, X, cir):
        """Sh3Converter::toBinaryMatrix(si64Matrix): X arithmetic shares [3][2][rows][cols] -> binary shares of the
        same values; cir = the flat adder circuit (harness.library_circuit("a2b", 64 * cols))."""
        X = np.ascontiguousarray(X, dtype=np.int64)
        rows = X.shape[2]
        x0, x1 = np.empty_like(X), np.empty_like(X)
        lib.orc_conv_a2b_inputs(self.h, ptr(X), X[0, 0].size, ptr(x0), ptr(x1))
        outs, _ = bin_eval(self, cir, rows, [x0, x1])
        return outs[0]

    def conv_bit_injection(self, B, bits):
        """Sh3Converter::bitInjection: B binary shares [3][2][rows][words] -> arithmetic shares [3][2][rows][bits]"""
        B = np.ascontiguousarray(B, dtype=np.int64)
        rows, words = B.shape[2], B.shape[3]
        Y = np.empty((3, 2, rows, bits), dtype=np.int64)
        lib.orc_conv_bit_injection(self.h, ptr(B), rows, words, bits, ptr(Y))
        return Y


def _dims(A, B, mode):
    assert A.dtype == np.int64 and B.dtype == np.int64 and A.flags.c_contiguous and B.flags.c_contiguous
    M, K = A.shape[2], A.shape[3]
    if mode == 1:
        assert A.shape == B.shape
        return M, 1, A.shape[3]
    assert B.shape[2] == K
    return M, K, B.shape[3]


def reveal(shares, party=0, binary=False):
    shares = np.ascontiguousarray(shares, dtype=np.int64)
    n = shares[0, 0].size
    out = np.empty(shares.shape[2:], dtype=np.int64)
    lib.orc_reveal(ptr(shares), n, party, int(binary), ptr(out))
    return out


def plain_mul(A, B, mode=0, nthreads=1):
    A = np.ascontiguousarray(A, dtype=np.int64)
    B = np.ascontiguousarray(B, dtype=np.int64)
    if mode == 1:
        M, K, N = A.shape[0], 1, A.shape[1]
    else:
        M, K, N = A.shape[0], A.shape[1], B.shape[1]
    out = np.empty((M, N), dtype=np.int64)
    lib.orc_plain_mul(ptr(A), ptr(B), ptr(out), M, K, N, mode, nthreads)
    return out


def cross_term(A0, A1, B0, B1, mode=0, nthreads=1):
    A0, A1, B0, B1 = (np.ascontiguousarray(x, dtype=np.int64) for x in (A0, A1, B0, B1))
    if mode == 1:
        M, K, N = A0.shape[0], 1, A0.shape[1]
    else:
        M, K, N = A0.shape[0], A0.shape[1], B0.shape[1]
    out = np.empty((M, N), dtype=np.int64)
    lib.orc_cross_term(ptr(A0), ptr(A1), ptr(B0), ptr(B1), ptr(out), M, K, N, mode, nthreads)
    return out


def bit_transpose(inp, rows, cols, in_stride, out_stride):
    inp = np.ascontiguousarray(inp, dtype=np.uint8)
    out = np.zeros(cols * out_stride, dtype=np.uint8)
    lib.orc_bit_transpose(ptr(inp), rows, cols, in_stride, ptr(out), out_stride)
    return out


class Circuit(C.Structure):
    _fields_ = [("wire_count", C.c_uint32), ("gate_count", C.c_uint32), ("gates", _p),
                ("level_count", C.c_uint32), ("level_gates", _p),
                ("num_inputs", C.c_uint32), ("input_first", _p), ("input_bits", _p),
                ("num_outputs", C.c_uint32), ("output_off", _p), ("output_bits", _p),
                ("output_wires", _p), ("output_invert", _p)]


def bin_eval(session, cir, width, inputs):
    """Run the oracle's Sh3BinaryEvaluator restatement.  cir: flat dict (harness.library_circuit
    layout); inputs: list of share arrays [3][2][width][words].  Returns (outputs, wire memory)."""
    c = Circuit()
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
    rb = lib.orc_bin_row_bytes(width)
    mem = np.zeros((3, 2, cir["wire_count"], rb), dtype=np.uint8)
    lib.orc_bin_eval(session.h, C.byref(c), width, in_ptrs, out_ptrs, ptr(mem))
    return outs, mem
