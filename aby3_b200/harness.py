"""ctypes binding of the sh3 facade's C harness (aby3_b200/sh3/harness.cpp, libsh3.so):
three in-process parties, each a worker thread with its own Sh3Runtime /
Sh3Encryptor / Sh3Evaluator and device stream.  Used by tests and bench.py."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsh3.so")


class Sh3Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise Sh3Error("libsh3.so is missing (%s); build it with `python -m aby3_b200.build` -- there is no CPU fallback" % LIB_PATH)
    C.CDLL(os.path.join(_HERE, "libaby3cu.so"), mode=C.RTLD_GLOBAL)
    l = C.CDLL(LIB_PATH)
    p, u64, i64, i32 = C.c_void_p, C.c_uint64, C.c_int64, C.c_int
    l.sh3h_last_error.restype = C.c_char_p
    l.sh3h_create.restype = p
    l.sh3h_create.argtypes = [i32, i32, i32, C.c_char_p, C.c_char_p]
    l.sh3h_create_nccl.restype = p
    l.sh3h_create_nccl.argtypes = [i32, i32, i32, C.c_char_p, C.c_char_p]
    l.sh3h_create_shared_stream.restype = p
    l.sh3h_create_shared_stream.argtypes = [i32, C.c_char_p, C.c_char_p]
    l.sh3h_destroy.argtypes = [p]
    l.sh3h_set_disable_randomization.argtypes = [p, i32]
    l.sh3h_set_gemm_algo.argtypes = [p, i32]
    l.sh3h_set_open_blocks.argtypes = [p, u64]
    l.sh3h_cursors.argtypes = [p, i32, p]
    l.sh3h_plain_create.argtypes = [p, i32, u64, u64, C.POINTER(p)]
    l.sh3h_plain_touch.argtypes = [p, i32, i32]
    l.sh3h_plain_prefetch.argtypes = [p, i32, i32]
    l.sh3h_reveal_plain_async.argtypes = [p, i32, i32, i32]
    l.sh3h_plain_wait.argtypes = [p, i32, i32]
    l.sh3h_share.argtypes = [p, i32, i32, u64, u64, i32, u64]
    l.sh3h_set_shares.argtypes = [p, p, u64, u64, i32, u64]
    l.sh3h_get_shares.argtypes = [p, i32, i32, p]
    l.sh3h_shape.argtypes = [p, i32, i32, C.POINTER(u64), C.POINTER(u64)]
    l.sh3h_free.argtypes = [p, i32]
    l.sh3h_mul.argtypes = [p, i32, i32, i64, i32]
    l.sh3h_addsub.argtypes = [p, i32, i32, i32]
    l.sh3h_cipher_gt.argtypes = [p, i32, i32]
    l.sh3h_max_min_split.argtypes = [p, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    l.sh3h_odd_even_merge.argtypes = [p, i32, i32]
    l.sh3h_piecewise.argtypes = [p, i32, p, i32, p, p, p, p, u64]
    l.sh3h_mul_bit.argtypes = [p, i32, i32, i32, i64]
    l.sh3h_share_packed.argtypes = [p, i32, i32, u64, u64, i32, p]
    l.sh3h_reveal_packed.argtypes = [p, i32, i32, p]
    l.sh3h_conv_init.argtypes = [p]
    l.sh3h_conv_a2b.argtypes = [p, i32, u64]
    l.sh3h_conv_bit_injection.argtypes = [p, i32, i32]
    l.sh3h_conv_packed_roundtrip.argtypes = [p, i32, p, C.POINTER(u64)]
    l.sh3h_reveal.argtypes = [p, i32, i32, i32, p]
    l.sh3h_reveal_plain.argtypes = [p, i32, i32, i32]
    l.sh3h_trunc_tuple.argtypes = [p, i32, u64, u64, u64, p, p, p]
    l.sh3h_bin_eval.argtypes = [p, p, C.c_uint32, C.c_uint32, p, C.c_uint32, p, p, C.c_uint32, p, p, p, p, C.c_uint32, p, p]
    l.sh3h_bin_eval_check.argtypes = [p, p, C.c_uint32, C.c_uint32, p, C.c_uint32, p, p, C.c_uint32, p, p, p, p, C.c_uint32, p, i64, i32, C.POINTER(u64)]
    l.sh3h_linreg.argtypes = [p, i32, i32, i32, p, u64, u64, C.c_double]
    l.sh3h_logreg.argtypes = l.sh3h_linreg.argtypes
    l.sh3h_linreg_graph.argtypes = l.sh3h_linreg.argtypes
    l.sh3h_linreg_fused.argtypes = l.sh3h_linreg.argtypes
    l.sh3h_bin_eval_packed.argtypes = l.sh3h_bin_eval.argtypes
    l.sh3h_timer_begin.argtypes = [p]
    l.sh3h_timer_end.argtypes = [p, C.POINTER(C.c_float)]
    l.sh3h_sync.argtypes = [p]
    l.sh3h_launch_count.restype = u64
    l.sh3h_launch_count.argtypes = [p]
    l.sh3h_trim.argtypes = [p]
    l.sh3h_pool_stats.argtypes = [p, p]
    l.sh3h_guard_selftest.argtypes = [p, u64, u64]
    l.sh3h_device_ptrs.argtypes = [p, i32, p]
    l.sh3h_alloc_shares.argtypes = [p, u64, u64]
    l.sh3h_party_stream.restype = p
    l.sh3h_party_stream.argtypes = [p, i32]
    l.sh3h_bytes_sent.restype = u64
    l.sh3h_bytes_sent.argtypes = [p]
    return l


lib = _load()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def to_block(hi, lo):
    return int(lo & (2**64 - 1)).to_bytes(8, "little") + int(hi & (2**64 - 1)).to_bytes(8, "little")


def default_seeds():
    """aby3_tests/Sh3EvaluatorTests.cpp:41-47"""
    enc = b"".join(to_block(0, i) + to_block(0, (i + 1) % 3) for i in range(3))
    ev = b"".join(to_block(1, i) + to_block(1, (i + 1) % 3) for i in range(3))
    return enc, ev


class Session:
    """Three parties on devices (d0, d1, d2) -- all 0 by default (co-located)."""

    def __init__(self, devices=(0, 0, 0), enc_seeds=None, eval_seeds=None, transport="local"):
        """transport "local": in-process hand-off (same GPU: D2D, other GPU: NVLink peer copy);
        "nccl": three different GPUs, ncclSend/ncclRecv over NVLink."""
        e, v = default_seeds()
        if transport == "shared_stream":
            # co-located parties on one stream: ordering by enqueue order, no events per message
            assert devices[0] == devices[1] == devices[2]
            self.h = lib.sh3h_create_shared_stream(devices[0], enc_seeds or e, eval_seeds or v)
        else:
            create = lib.sh3h_create_nccl if transport == "nccl" else lib.sh3h_create
            self.h = create(devices[0], devices[1], devices[2], enc_seeds or e, eval_seeds or v)
        if not self.h:
            raise Sh3Error(lib.sh3h_last_error().decode())

    def _chk(self, rc):
        if rc != 0:
            raise Sh3Error(lib.sh3h_last_error().decode())

    def _id(self, rc):
        if rc < 0:
            raise Sh3Error(lib.sh3h_last_error().decode())
        return rc

    def close(self):
        if self.h:
            lib.sh3h_destroy(self.h)
            self.h = None

    def disable_randomization(self, on=True):
        lib.sh3h_set_disable_randomization(self.h, int(on))

    def set_open_blocks(self, blocks):
        """row blocks in which the opened xy - r of a truncating matrix product travels (1 = one message)"""
        lib.sh3h_set_open_blocks(self.h, int(blocks))

    def set_gemm_algo(self, algo):
        lib.sh3h_set_gemm_algo(self.h, int(algo))

    def cursors(self, party):
        c = np.zeros(6, dtype=np.uint64)
        lib.sh3h_cursors(self.h, party, _ptr(c))
        return c

    def plain(self, owner, rows, cols):
        """Plaintext matrix held by `owner` in page-locked host memory.
        Returns (handle, numpy view to fill)."""
        hp = C.c_void_p()
        hid = self._id(lib.sh3h_plain_create(self.h, owner, rows, cols, C.byref(hp)))
        buf = (C.c_int64 * (rows * cols)).from_address(hp.value)
        return hid, np.frombuffer(buf, dtype=np.int64).reshape(rows, cols)

    def plain_touch(self, owner, hid):
        self._chk(lib.sh3h_plain_touch(self.h, owner, hid))

    def plain_prefetch(self, owner, hid):
        """Start uploading a plaintext matrix on the owner's copy stream (overlaps queued kernels); the next device
        use of the matrix waits for it.  Do not write the numpy view until the data has been shared."""
        self._chk(lib.sh3h_plain_prefetch(self.h, owner, hid))

    def reveal_plain_async(self, hid, party, plain_id):
        """revealAll; `party`'s result is downloaded into its plaintext matrix on the copy stream without blocking.
        Call plain_wait before reading the numpy view."""
        self._chk(lib.sh3h_reveal_plain_async(self.h, hid, party, plain_id))

    def plain_wait(self, party, plain_id):
        self._chk(lib.sh3h_plain_wait(self.h, party, plain_id))

    def share_int(self, owner, values):
        values = np.ascontiguousarray(values, dtype=np.int64)
        if values.ndim == 1:
            values = values.reshape(-1, 1)
        pid, view = self.plain(owner, values.shape[0], values.shape[1])
        view[...] = values
        hid = self._id(lib.sh3h_share(self.h, owner, pid, values.shape[0], values.shape[1], 0, 0))
        self.free(pid)
        return hid

    def share_plain(self, owner, pid, rows, cols):
        return self._id(lib.sh3h_share(self.h, owner, pid, rows, cols, 0, 0))

    def share_bin(self, owner, values, bit_count):
        values = np.ascontiguousarray(values, dtype=np.int64)
        if values.ndim == 1:
            values = values.reshape(-1, 1)
        assert values.shape[1] == (bit_count + 63) // 64
        pid, view = self.plain(owner, values.shape[0], values.shape[1])
        view[...] = values
        hid = self._id(lib.sh3h_share(self.h, owner, pid, values.shape[0], values.shape[1], 1, bit_count))
        self.free(pid)
        return hid

    def set_shares(self, shares, binary=False, bit_count=0):
        shares = np.ascontiguousarray(shares, dtype=np.int64)
        rows, cols = shares.shape[2], shares.shape[3]
        return self._id(lib.sh3h_set_shares(self.h, _ptr(shares), rows, cols if not binary else 0, int(binary), bit_count))

    def shape(self, hid, binary=False):
        r, c = C.c_uint64(), C.c_uint64()
        self._chk(lib.sh3h_shape(self.h, hid, int(binary), C.byref(r), C.byref(c)))
        return r.value, c.value

    def get_shares(self, hid, binary=False):
        r, c = self.shape(hid, binary)
        out = np.empty((3, 2, r, c), dtype=np.int64)
        self._chk(lib.sh3h_get_shares(self.h, hid, int(binary), _ptr(out)))
        return out

    def free(self, hid):
        self._chk(lib.sh3h_free(self.h, hid))

    def mul(self, a, b, shift=None, out=0):
        return self._id(lib.sh3h_mul(self.h, a, b, -1 if shift is None else int(shift), out))

    def mul_bit(self, a, b):
        """c = b * a for a one-bit binary sharing b (asyncMul(si64Matrix, sbMatrix))."""
        return self._id(lib.sh3h_mul_bit(self.h, a, b, 0, 0))

    def mul_bit_pub(self, a_const, b):
        """c = b * a for a public constant a (asyncMul(i64, sbMatrix))."""
        return self._id(lib.sh3h_mul_bit(self.h, 0, b, 1, int(a_const)))

    def piecewise(self, x, thresholds, coefficients, D=16):
        """Sh3Piecewise::eval.  coefficients: one list per region, constant term first; python ints are
        integer coefficients, floats fixed-point ones (aby3ML::logisticFunc uses
        thresholds [-0.5, 0.5], coefficients [[], [0.5, 1], [1]])."""
        th = np.asarray(thresholds, dtype=np.float64)
        counts = np.asarray([len(c) for c in coefficients], dtype=np.int32)
        flat = [v for c in coefficients for v in c]
        is_int = np.asarray([isinstance(v, (int, np.integer)) for v in flat] or [0], dtype=np.int32)
        ints = np.asarray([int(v) if isinstance(v, (int, np.integer)) else 0 for v in flat] or [0], dtype=np.int64)
        dbl = np.asarray([float(v) for v in flat] or [0.0], dtype=np.float64)
        return self._id(lib.sh3h_piecewise(self.h, x, _ptr(th), len(th), _ptr(counts), _ptr(is_int), _ptr(ints), _ptr(dbl), D))

    def share_packed(self, owner, values, use_task=False):
        """Sh3Encryptor::localPackedBinary / remotePackedBinary: rows secrets of 64 * cols bits, bit-sliced.
        Returns (handle, share planes [3][2][bits][simd])."""
        values = np.ascontiguousarray(values, dtype=np.int64)
        rows, cols = values.shape
        pid, view = self.plain(owner, rows, cols)
        view[...] = values
        sh = np.zeros((3, 2, 64 * cols, (rows + 63) // 64), dtype=np.int64)
        hid = self._id(lib.sh3h_share_packed(self.h, owner, pid, rows, cols, int(use_task), _ptr(sh)))
        self.free(pid)
        return hid, sh

    def reveal_packed(self, hid, party, rows, cols):
        out = np.empty((rows, cols), dtype=np.int64)
        self._chk(lib.sh3h_reveal_packed(self.h, hid, party, _ptr(out)))
        return out

    def conv_init(self):
        """Sh3Converter::init(rt, eval.mShareGen) on every party"""
        self._chk(lib.sh3h_conv_init(self.h))

    def conv_a2b(self, x, bits=0):
        """Sh3Converter::toBinaryMatrix(si64Matrix): arithmetic sharing -> binary sharing (64 bits per word, or a
        pre-sized destination of `bits` bits per row)"""
        return self._id(lib.sh3h_conv_a2b(self.h, x, int(bits)))

    def conv_bit_injection(self, b, two_rounds=False):
        """Sh3Converter::bitInjection: binary sharing (rows x bits) -> arithmetic sharing, one element per bit"""
        return self._id(lib.sh3h_conv_bit_injection(self.h, b, int(two_rounds)))

    def conv_packed_roundtrip(self, b, bit_count):
        """toPackedBin then toBinaryMatrix(sPackedBin); returns (new handle, packed planes [3][2][bit_count][simd])"""
        rows, _ = self.shape(b, binary=True)
        simd = (rows + 63) // 64
        packed = np.zeros((3, 2, bit_count, simd), dtype=np.int64)
        sw = C.c_uint64(0)
        hid = self._id(lib.sh3h_conv_packed_roundtrip(self.h, b, _ptr(packed), C.byref(sw)))
        assert sw.value == simd
        return hid, packed

    def cipher_gt(self, a, b):
        """aby3-Basic cipher_gt: one-bit binary sharing of (a > b) for arithmetic sharings a, b."""
        return self._id(lib.sh3h_cipher_gt(self.h, a, b))

    def max_min_split(self, a, b):
        """aby3-Basic bool_cipher_max_min_split on binary sharings of 64-bit values -> (max, min)."""
        mx, mn = C.c_int(0), C.c_int(0)
        self._chk(lib.sh3h_max_min_split(self.h, a, b, C.byref(mx), C.byref(mn)))
        return mx.value, mn.value

    def odd_even_merge(self, a, b):
        """aby3-Basic odd_even_merge of two sorted binary sharings."""
        return self._id(lib.sh3h_odd_even_merge(self.h, a, b))

    def add(self, a, b):
        return self._id(lib.sh3h_addsub(self.h, a, b, 0))

    def sub(self, a, b):
        return self._id(lib.sh3h_addsub(self.h, a, b, 1))

    def reveal(self, hid, party=0, binary=False, out=None):
        r, c = self.shape(hid, binary)
        if out is None:
            out = np.empty((r, c), dtype=np.int64)
        self._chk(lib.sh3h_reveal(self.h, hid, int(binary), party, _ptr(out)))
        return out

    def reveal_plain(self, hid, party, plain_id):
        """revealAll; `party`'s result lands in its page-locked plaintext matrix `plain_id`
        (the numpy view returned by plain() stays valid when the shape matches)."""
        self._chk(lib.sh3h_reveal_plain(self.h, hid, party, plain_id))

    def trunc_tuple(self, party, rows, cols, d):
        n = rows * cols
        R, T0, T1 = (np.empty(n, dtype=np.int64) for _ in range(3))
        self._chk(lib.sh3h_trunc_tuple(self.h, party, rows, cols, d, _ptr(R), _ptr(T0), _ptr(T1)))
        return R, T0, T1

    def bin_eval(self, cir, input_ids, packed=False):
        """cir: flat circuit dict (library_circuit). Returns output handles.  packed=True routes
        inputs/outputs through sPackedBin (setInput/getOutput bit-sliced forms)."""
        ins = np.asarray(input_ids, dtype=np.int32)
        outs = np.zeros(len(cir["output_bits"]), dtype=np.int32)
        inv = cir.get("output_invert")
        fn = lib.sh3h_bin_eval_packed if packed else lib.sh3h_bin_eval
        self._chk(fn(
            self.h, _ptr(cir["gates"]), len(cir["gates"]) // 4, cir["wire_count"], _ptr(cir["level_gates"]),
            len(cir["level_gates"]), _ptr(cir["input_first"]), _ptr(cir["input_bits"]), len(cir["input_bits"]),
            _ptr(cir["output_off"]), _ptr(cir["output_bits"]), _ptr(cir["output_wires"]),
            _ptr(inv) if inv is not None else None, len(cir["output_bits"]), _ptr(ins), _ptr(outs)))
        return [int(x) for x in outs]

    def bin_eval_check(self, cir, input_ids, tamper_wire=-1, tamper_party=0):
        """Evaluate `cir`, then the device-side shadow check of every gate on the reconstructed wires
        (Sh3BinaryEvaluator::enableDebug / validateMemory).  Returns the number of disagreeing instance-gates summed over
        the three parties; tamper_wire >= 0 injects a one-bit fault into that wire at `tamper_party` first."""
        ins = np.asarray(input_ids, dtype=np.int32)
        inv = cir.get("output_invert")
        bad = C.c_uint64(0)
        self._chk(lib.sh3h_bin_eval_check(
            self.h, _ptr(cir["gates"]), len(cir["gates"]) // 4, cir["wire_count"], _ptr(cir["level_gates"]),
            len(cir["level_gates"]), _ptr(cir["input_first"]), _ptr(cir["input_bits"]), len(cir["input_bits"]),
            _ptr(cir["output_off"]), _ptr(cir["output_bits"]), _ptr(cir["output_wires"]),
            _ptr(inv) if inv is not None else None, len(cir["output_bits"]), _ptr(ins), int(tamper_wire), int(tamper_party), C.byref(bad)))
        return int(bad.value)

    def linreg(self, X, Y, w, batch_idx, iters, batch, lr):
        """aby3-ML SGD_Linear (ml/Regression.h) on sf64<D16> shares; w is updated in place."""
        batch_idx = np.ascontiguousarray(batch_idx, dtype=np.uint64)
        assert batch_idx.size == iters * batch
        self._chk(lib.sh3h_linreg(self.h, X, Y, w, _ptr(batch_idx), iters, batch, float(lr)))

    def linreg_graph(self, X, Y, w, batch_idx, iters, batch, lr):
        """SGD_Linear for co-located parties, one CUDA-graph launch per iteration (ml/SgdGraph.h); same result as linreg."""
        batch_idx = np.ascontiguousarray(batch_idx, dtype=np.uint64)
        assert batch_idx.size == iters * batch
        self._chk(lib.sh3h_linreg_graph(self.h, X, Y, w, _ptr(batch_idx), iters, batch, float(lr)))

    def linreg_fused(self, X, Y, w, batch_idx, iters, batch, lr):
        """SGD_Linear for co-located parties, the whole run as one persistent kernel (csrc/sgd_fused.cu); same result as linreg."""
        batch_idx = np.ascontiguousarray(batch_idx, dtype=np.uint64)
        assert batch_idx.size == iters * batch
        self._chk(lib.sh3h_linreg_fused(self.h, X, Y, w, _ptr(batch_idx), iters, batch, float(lr)))

    def logreg(self, X, Y, w, batch_idx, iters, batch, lr):
        """aby3-ML SGD_Logistic (ml/Regression.h) on sf64<D16> shares; w is updated in place."""
        batch_idx = np.ascontiguousarray(batch_idx, dtype=np.uint64)
        assert batch_idx.size == iters * batch
        self._chk(lib.sh3h_logreg(self.h, X, Y, w, _ptr(batch_idx), iters, batch, float(lr)))

    def timer_begin(self):
        self._chk(lib.sh3h_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_float(0)
        self._chk(lib.sh3h_timer_end(self.h, C.byref(ms)))
        return float(ms.value)

    def sync(self):
        self._chk(lib.sh3h_sync(self.h))

    @property
    def launches(self):
        return int(lib.sh3h_launch_count(self.h))

    def trim(self):
        """Return the cached device blocks of the three buffer pools to the driver (call between workloads
        whose buffer sizes differ, so that one workload's cache is not the next one's memory pressure)."""
        self._chk(lib.sh3h_trim(self.h))

    @property
    def pool_stats(self):
        """(driver mallocs, bytes, driver frees) behind the three parties' buffer pools"""
        a = np.zeros(3, dtype=np.uint64)
        lib.sh3h_pool_stats(self.h, _ptr(a))
        return tuple(int(x) for x in a)

    def device_ptrs(self, hid):
        """device addresses [party][plane] of an arithmetic sharing (interop: e.g. wrap them as torch tensors)"""
        arr = (C.c_void_p * 6)()
        self._chk(lib.sh3h_device_ptrs(self.h, hid, arr))
        return [[int(arr[2 * i] or 0), int(arr[2 * i + 1] or 0)] for i in range(3)]

    def alloc_shares(self, rows, cols):
        """an arithmetic sharing with device planes allocated but not written (filled through device_ptrs)"""
        return self._id(lib.sh3h_alloc_shares(self.h, rows, cols))

    def party_stream(self, party):
        """the party's cudaStream_t as an integer"""
        return int(lib.sh3h_party_stream(self.h, party) or 0)

    def guard_selftest(self, nbytes, overrun):
        """ABY3_POOL_GUARD=1 only: write `overrun` bytes past a pool block; 1 = the guard caught it, 0 = clean, -1 = guard off"""
        return int(lib.sh3h_guard_selftest(self.h, int(nbytes), int(overrun)))

    @property
    def bytes_sent(self):
        return int(lib.sh3h_bytes_sent(self.h))


lib.sh3h_circuit_build.restype = C.c_void_p
lib.sh3h_circuit_build.argtypes = [C.c_char_p, C.c_uint32]
lib.sh3h_circuit_free.argtypes = [C.c_void_p]
lib.sh3h_circuit_sizes.argtypes = [C.c_void_p, C.c_void_p]
lib.sh3h_circuit_copy.argtypes = [C.c_void_p] + [C.c_void_p] * 8


def library_circuit(name, bits):
    """Flat description of a circuit from the facade's BetaLibrary (host only)."""
    h = lib.sh3h_circuit_build(name.encode(), bits)
    if not h:
        raise Sh3Error(lib.sh3h_last_error().decode())
    s = np.zeros(7, dtype=np.uint32)
    lib.sh3h_circuit_sizes(h, _ptr(s))
    ng, nw, nl, ni, no, now, nand = (int(x) for x in s)
    cir = {
        "wire_count": nw, "nonlinear": nand,
        "gates": np.zeros(4 * ng, dtype=np.uint32), "level_gates": np.zeros(nl, dtype=np.uint32),
        "input_first": np.zeros(ni, dtype=np.uint32), "input_bits": np.zeros(ni, dtype=np.uint32),
        "output_off": np.zeros(no, dtype=np.uint32), "output_bits": np.zeros(no, dtype=np.uint32),
        "output_wires": np.zeros(now, dtype=np.uint32), "output_invert": np.zeros(now, dtype=np.uint8),
    }
    lib.sh3h_circuit_copy(h, _ptr(cir["gates"]), _ptr(cir["level_gates"]), _ptr(cir["input_first"]), _ptr(cir["input_bits"]),
                          _ptr(cir["output_off"]), _ptr(cir["output_bits"]), _ptr(cir["output_wires"]), _ptr(cir["output_invert"]))
    lib.sh3h_circuit_free(h)
    return cir
