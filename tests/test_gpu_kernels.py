"""GPU parity tests, kernel by kernel, through the C ABI (include/aby3cu.h),
against the CPU oracle on the same keys and offsets.  Bit-exact: all integer."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as o
from aby3_b200 import abi

pytestmark = pytest.mark.gpu
U64 = np.uint64
P = C.c_void_p
lib = abi.lib

KEY_A = bytes(range(16))
KEY_B = bytes(range(100, 116))
SIZES = [1, 2, 3, 16, 511, 512, 513, 4097, (1 << 20) + 5]


def rnd(seed, n):
    return np.random.default_rng(seed).integers(-2**63, 2**63, n, dtype=np.int64)


def test_aes_ctr_fill_matches_oracle(ctx):
    for n in SIZES:
        for e0 in (0, 4, 7, 2**33 + 1):
            buf = ctx.alloc(8 * n + 16)
            abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, KEY_A, 8 * e0, buf.p, 8 * n))
            got = ctx.download(buf, n, U64)
            assert np.array_equal(got, o.stream_u64(KEY_A, e0, n)), (n, e0)
            buf.free()


def test_aes_ctr_fill_unaligned_destination(ctx):
    n = 1001
    buf = ctx.alloc(8 * n + 64)
    abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, KEY_B, 8 * 3, buf.at(8), 8 * n))
    got = ctx.download(buf, n, U64, byte_off=8)
    assert np.array_equal(got, o.stream_u64(KEY_B, 3, n))


def test_aes_fips197_on_device(ctx):
    """AES_key(toBlock(c)) for the FIPS-197 C.1 key equals the software AES of the
    same counter block -- pins the device T-tables to the standard."""
    key = bytes.fromhex("000102030405060708090a0b0c0d0e0f")
    buf = ctx.alloc(16 * 4)
    abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, key, 16 * 0x0766554433221100, buf.p, 64))
    got = ctx.download(buf, 64, np.uint8)
    for i in range(4):
        ref = np.zeros(16, np.uint8)
        pt = (0x0766554433221100 + i).to_bytes(8, "little") + bytes(8)
        o.lib.orc_aes128_encrypt(key, pt, o.ptr(ref), 1)
        assert got[16 * i:16 * i + 16].tobytes() == ref.tobytes()


@pytest.mark.parametrize("binary", [0, 1])
def test_zero_share(ctx, binary):
    for n in SIZES:
        for e0 in (0, 1, 511):
            add = rnd(n, n)
            d_add, d_out = ctx.upload(add), ctx.alloc(8 * n + 16)
            abi.check(lib.aby3cu_zero_share(ctx.h, KEY_A, KEY_B, e0, d_add.p, d_out.p, n, binary))
            got = ctx.download(d_out, n, U64)
            a, b = o.stream_u64(KEY_A, e0, n), o.stream_u64(KEY_B, e0, n)
            exp = (add.view(U64) ^ a ^ b) if binary else (add.view(U64) + (a - b))
            assert np.array_equal(got, exp), (n, e0)
            abi.check(lib.aby3cu_zero_share(ctx.h, KEY_A, KEY_B, e0, None, d_out.p, n, binary))
            got = ctx.download(d_out, n, U64)
            assert np.array_equal(got, (a ^ b) if binary else (a - b))
            d_add.free(); d_out.free()


def test_mul_hadamard(ctx):
    for n in SIZES:
        a0, a1, b0, b1 = (rnd(s + n, n) for s in range(4))
        d = [ctx.upload(x) for x in (a0, a1, b0, b1)]
        out = ctx.alloc(8 * n + 16)
        e0 = 77
        abi.check(lib.aby3cu_mul_hadamard(ctx.h, d[0].p, d[1].p, d[2].p, d[3].p, KEY_A, KEY_B, e0, out.p, n))
        got = ctx.download(out, n, U64)
        cross = o.cross_term(a0.reshape(1, n), a1.reshape(1, n), b0.reshape(1, n), b1.reshape(1, n), mode=1).reshape(n).view(U64)
        z = o.stream_u64(KEY_A, e0, n) - o.stream_u64(KEY_B, e0, n)
        assert np.array_equal(got, cross + z)
        abi.check(lib.aby3cu_mul_hadamard(ctx.h, d[0].p, d[1].p, d[2].p, d[3].p, None, None, 0, out.p, n))
        assert np.array_equal(ctx.download(out, n, U64), cross)
        for x in d + [out]:
            x.free()


@pytest.mark.parametrize("rand", [True, False])
def test_trunc_tuple_and_hadamard_trunc(ctx, rand):
    for n in [1, 2, 513, 100003]:
        for d, en, ep in [(8, 4, 4), (16, 5, 4), (33, 4, 9)]:
            kn, kp = (KEY_A, KEY_B) if rand else (None, None)
            R, NEG, T0, T1 = (ctx.alloc(8 * n + 16) for _ in range(4))
            abi.check(lib.aby3cu_trunc_tuple(ctx.h, kn, en, kp, ep, d, R.p, NEG.p, T0.p, T1.p, n))
            r, neg, t0, t1 = (ctx.download(x, n, np.int64) for x in (R, NEG, T0, T1))
            s0 = o.stream_u64(KEY_A, en, n).view(np.int64) if rand else np.zeros(n, np.int64)
            s1 = o.stream_u64(KEY_B, ep, n).view(np.int64) if rand else np.zeros(n, np.int64)
            assert np.array_equal(r, s0 >> 2)
            assert np.array_equal(neg.view(U64), U64(0) - (s0 >> 2).view(U64))
            assert np.array_equal(t0, s0 >> (d + 2))
            assert np.array_equal(t1, s1 >> (d + 2))
            a0, a1, b0, b1 = (rnd(s + 11 * n, n) for s in range(4))
            dd = [ctx.upload(x) for x in (a0, a1, b0, b1)]
            V = ctx.alloc(8 * n + 16)
            abi.check(lib.aby3cu_mul_hadamard_trunc(ctx.h, dd[0].p, dd[1].p, dd[2].p, dd[3].p, kn, en, kp, ep, d,
                                                    V.p, T0.p, T1.p, n))
            v = ctx.download(V, n, U64)
            cross = o.cross_term(a0.reshape(1, n), a1.reshape(1, n), b0.reshape(1, n), b1.reshape(1, n), mode=1).reshape(n).view(U64)
            assert np.array_equal(v, cross - (s0 >> 2).view(U64))
            assert np.array_equal(ctx.download(T0, n, np.int64), s0 >> (d + 2))
            assert np.array_equal(ctx.download(T1, n, np.int64), s1 >> (d + 2))
            for x in dd + [R, NEG, T0, T1, V]:
                x.free()


def test_trunc_finish(ctx):
    for n in [1, 2, 513, 100003]:
        for shift in (8, 16, 33):
            s0, s1, s2, c = (rnd(s + n, n) for s in range(4))
            d = [ctx.upload(x) for x in (s0, s1, s2, c)]
            abi.check(lib.aby3cu_trunc_finish(ctx.h, d[0].p, d[1].p, d[2].p, d[3].p, n, shift))
            got = ctx.download(d[3], n, U64)
            tot = (s0.view(U64) + s1.view(U64) + s2.view(U64)).view(np.int64)
            assert np.array_equal(got, c.view(U64) + (tot >> shift).view(U64))
            for x in d:
                x.free()


def test_share_ops_and_combine(ctx):
    n = 70001
    x, y, z = rnd(1, n), rnd(2, n), rnd(3, n)
    dx, dy, dz, out = ctx.upload(x), ctx.upload(y), ctx.upload(z), ctx.alloc(8 * n)
    for op, f in [(abi.OP_ADD, lambda a, b: a + b), (abi.OP_SUB, lambda a, b: a - b), (abi.OP_XOR, lambda a, b: a ^ b)]:
        abi.check(lib.aby3cu_share_op(ctx.h, op, dx.p, dy.p, out.p, n))
        assert np.array_equal(ctx.download(out, n, U64), f(x.view(U64), y.view(U64)))
    abi.check(lib.aby3cu_combine3(ctx.h, abi.OP_ADD, dx.p, dy.p, dz.p, out.p, n))
    assert np.array_equal(ctx.download(out, n, U64), x.view(U64) + y.view(U64) + z.view(U64))
    abi.check(lib.aby3cu_combine3(ctx.h, abi.OP_XOR, dx.p, dy.p, dz.p, out.p, n))
    assert np.array_equal(ctx.download(out, n, U64), x.view(U64) ^ y.view(U64) ^ z.view(U64))


def test_transpose_and_gather(ctx):
    for r, c in [(1, 1), (128, 1024), (33, 65), (1000, 3)]:
        m = rnd(r * c, r * c).reshape(r, c)
        dm, out = ctx.upload(m), ctx.alloc(8 * r * c)
        abi.check(lib.aby3cu_transpose_i64(ctx.h, dm.p, r, c, out.p))
        assert np.array_equal(ctx.download(out, (c, r)), m.T)
    m = rnd(5, 500 * 37).reshape(500, 37)
    idx = np.random.default_rng(0).integers(0, 500, 128).astype(np.uint64)
    dm, di, out = ctx.upload(m), ctx.upload(idx), ctx.alloc(8 * 128 * 37)
    abi.check(lib.aby3cu_gather_rows(ctx.h, dm.p, 37, di.p, 128, out.p))
    assert np.array_equal(ctx.download(out, (128, 37)), m[idx.astype(np.int64)])


GEMM_SHAPES = [(1, 1, 1), (10, 10, 10), (128, 64, 16), (129, 65, 17), (200, 300, 70), (128, 1024, 1),
               (1024, 128, 1), (77, 513, 3), (64, 64, 8), (5, 2000, 9), (256, 256, 256)]


@pytest.mark.parametrize("algo", [abi.GEMM_IMAD, abi.GEMM_AUTO])
def test_gemm_cross(ctx, algo):
    for (M, K, N) in GEMM_SHAPES:
        a0, a1 = rnd(1, M * K).reshape(M, K), rnd(2, M * K).reshape(M, K)
        b0, b1 = rnd(3, K * N).reshape(K, N), rnd(4, K * N).reshape(K, N)
        c0 = rnd(5, M * N).reshape(M, N)
        d = [ctx.upload(x) for x in (a0, a1, b0, b1)]
        exp = o.cross_term(a0, a1, b0, b1).view(U64)
        dc = ctx.upload(c0)
        abi.check(lib.aby3cu_gemm_cross(ctx.h, algo, d[0].p, d[1].p, d[2].p, d[3].p, M, K, N, dc.p, 1))
        assert np.array_equal(ctx.download(dc, (M, N), U64), exp + c0.view(U64)), (M, K, N, "acc")
        abi.check(lib.aby3cu_gemm_cross(ctx.h, algo, d[0].p, d[1].p, d[2].p, d[3].p, M, K, N, dc.p, 0))
        assert np.array_equal(ctx.download(dc, (M, N), U64), exp), (M, K, N)
        for x in d + [dc]:
            x.free()


def test_bit_transpose_matches_oracle(ctx):
    rng = np.random.default_rng(8)
    for rows, cols in [(1, 1), (7, 64), (100, 13), (256, 256), (65, 130), (5000, 64), (64, 5000), (3000, 128), (1025, 33)]:
        ins = ((cols + 31) // 32) * 4 + 8
        outs = ((rows + 31) // 32) * 4 + 4
        m = rng.integers(0, 256, rows * ins, dtype=np.uint8)
        d_in, d_out = ctx.upload(m), ctx.alloc(cols * outs)
        abi.check(lib.aby3cu_memset(ctx.h, d_out.p, 0, cols * outs))
        abi.check(lib.aby3cu_bit_transpose(ctx.h, d_in.p, rows, cols, ins, d_out.p, outs, None))
        got = ctx.download(d_out, cols * outs, np.uint8).reshape(cols, outs)
        exp = o.bit_transpose(m, rows, cols, ins, outs).reshape(cols, outs)
        nb = (rows + 7) // 8
        gb = np.unpackbits(got[:, :nb], axis=1, bitorder="little")[:, :rows]
        eb = np.unpackbits(exp[:, :nb], axis=1, bitorder="little")[:, :rows]
        assert np.array_equal(gb, eb), (rows, cols)
        d_in.free(); d_out.free()


def test_bit_transpose_round_trip_full_size(ctx):
    """size-independent property at a BASELINE-scale width: T(T(x)) == x."""
    width, bits = 1 << 22, 64
    x = rnd(3, width)
    rb = lib.aby3cu_bin_row_bytes(width)
    d_x, d_m, d_y = ctx.upload(x), ctx.alloc(bits * rb), ctx.alloc(8 * width)
    abi.check(lib.aby3cu_bit_transpose(ctx.h, d_x.p, width, bits, 8, d_m.p, rb, None))
    abi.check(lib.aby3cu_bit_transpose(ctx.h, d_m.p, bits, width, rb, d_y.p, 8, None))
    assert np.array_equal(ctx.download(d_y, width), x)


def test_two_plane_ops_and_multi_gather(ctx):
    """share_op2 / transpose_i64_2 / gather_rows_multi: both share planes (and X with Y) in one launch"""
    for n in (1, 2, 3, 513, 100003):
        x0, y0, x1, y1 = (rnd(50 + k, n) for k in range(4))
        d = [ctx.upload(a) for a in (x0, y0, x1, y1)]
        o0, o1 = ctx.alloc(8 * n), ctx.alloc(8 * n)
        for op, f in ((0, lambda a, b: a.view(U64) + b.view(U64)), (1, lambda a, b: a.view(U64) - b.view(U64)),
                      (2, lambda a, b: a.view(U64) ^ b.view(U64))):
            abi.check(lib.aby3cu_share_op2(ctx.h, op, d[0].p, d[1].p, o0.p, d[2].p, d[3].p, o1.p, n))
            assert np.array_equal(ctx.download(o0, n, U64), f(x0, y0))
            assert np.array_equal(ctx.download(o1, n, U64), f(x1, y1))
        # in place on the first operand (sf64Matrix::operator-=)
        abi.check(lib.aby3cu_share_op2(ctx.h, 1, d[0].p, d[1].p, d[0].p, d[2].p, d[3].p, d[2].p, n))
        assert np.array_equal(ctx.download(d[0], n, U64), x0.view(U64) - y0.view(U64))
        assert np.array_equal(ctx.download(d[2], n, U64), x1.view(U64) - y1.view(U64))
    for r, c in ((1, 1), (128, 1024), (33, 65), (1000, 7)):
        m0, m1 = rnd(60, r * c).reshape(r, c), rnd(61, r * c).reshape(r, c)
        d0, d1 = ctx.upload(m0), ctx.upload(m1)
        t0, t1 = ctx.alloc(8 * r * c), ctx.alloc(8 * r * c)
        abi.check(lib.aby3cu_transpose_i64_2(ctx.h, d0.p, d1.p, r, c, t0.p, t1.p))
        assert np.array_equal(ctx.download(t0, (c, r)), m0.T)
        assert np.array_equal(ctx.download(t1, (c, r)), m1.T)
    rows, cols = 500, 38
    X0, X1 = rnd(70, rows * cols).reshape(rows, cols), rnd(71, rows * cols).reshape(rows, cols)
    Y0, Y1 = rnd(72, rows).reshape(rows, 1), rnd(73, rows).reshape(rows, 1)
    idx = np.random.default_rng(74).integers(0, rows, 128).astype(np.uint64)
    srcs = [ctx.upload(a) for a in (X0, X1, Y0, Y1)]
    outs = [ctx.alloc(8 * 128 * cols), ctx.alloc(8 * 128 * cols), ctx.alloc(8 * 128), ctx.alloc(8 * 128)]
    di = ctx.upload(idx)
    in_p = (C.c_void_p * 4)(*[b.p for b in srcs])
    out_p = (C.c_void_p * 4)(*[b.p for b in outs])
    cols_a = (C.c_uint64 * 4)(cols, cols, 1, 1)
    abi.check(lib.aby3cu_gather_rows_multi(ctx.h, 4, in_p, cols_a, out_p, di.p, 128))
    for src, out, w in zip((X0, X1, Y0, Y1), outs, (cols, cols, 1, 1)):
        assert np.array_equal(ctx.download(out, (128, w)), src[idx.astype(np.int64)])
    # odd column count (no 16-byte rows) and fewer jobs
    Z = rnd(75, rows * 37).reshape(rows, 37)
    dz, oz = ctx.upload(Z), ctx.alloc(8 * 128 * 37)
    in1, out1, c1 = (C.c_void_p * 1)(dz.p), (C.c_void_p * 1)(oz.p), (C.c_uint64 * 1)(37)
    abi.check(lib.aby3cu_gather_rows_multi(ctx.h, 1, in1, c1, out1, di.p, 128))
    assert np.array_equal(ctx.download(oz, (128, 37)), Z[idx.astype(np.int64)])


def test_gemm_skinny_shapes(ctx):
    """GEMV-like shapes (N <= 8) of the CUDA-core path: odd K (scalar loads), K beyond one shared-memory chunk
    (accumulation across launches), every N bucket, unaligned A planes."""
    for (M, K, N) in [(1, 1, 1), (37, 513, 1), (300, 4100, 1), (64, 2048, 2), (33, 1500, 3), (129, 700, 5), (17, 300, 8), (5000, 512, 1)]:
        a0, a1 = rnd(11, M * K + 1), rnd(12, M * K + 1)
        b0, b1 = rnd(13, K * N).reshape(K, N), rnd(14, K * N).reshape(K, N)
        c0 = rnd(15, M * N).reshape(M, N)
        d = [ctx.upload(x) for x in (a0, a1, b0, b1)]
        for off in (0, 1):                       # off = 1: planes start 8 bytes off a 16-byte boundary
            A0, A1 = a0[off:off + M * K].reshape(M, K), a1[off:off + M * K].reshape(M, K)
            exp = o.cross_term(np.ascontiguousarray(A0), np.ascontiguousarray(A1), b0, b1).view(U64)
            dc = ctx.upload(c0)
            abi.check(lib.aby3cu_gemm_cross(ctx.h, abi.GEMM_IMAD, d[0].at(8 * off), d[1].at(8 * off), d[2].p, d[3].p, M, K, N, dc.p, 1))
            assert np.array_equal(ctx.download(dc, (M, N), U64), exp + c0.view(U64)), (M, K, N, off, "acc")
            abi.check(lib.aby3cu_gemm_cross(ctx.h, abi.GEMM_AUTO, d[0].at(8 * off), d[1].at(8 * off), d[2].p, d[3].p, M, K, N, dc.p, 0))
            assert np.array_equal(ctx.download(dc, (M, N), U64), exp), (M, K, N, off)
            dc.free()
        for x in d:
            x.free()


def test_graph_capture_and_iteration_relative_entry_points(ctx):
    """aby3cu_capture_begin/end + graph_launch with the *_at entry points: a captured (truncation pair, gather, counter
    increment) replayed k times produces what the plain entry points give at offsets base + it * stride."""
    n, stride, d, rows, cols, nb = 37, 50, 16, 40, 6, 8
    idx = np.random.default_rng(90).integers(0, rows, 5 * nb).astype(np.uint64)
    src = rnd(91, rows * cols).reshape(rows, cols)
    d_idx, d_src = ctx.upload(idx), ctx.upload(src)
    d_iter = ctx.upload(np.zeros(2, dtype=np.uint64))
    negr, rt0, rt1, out = ctx.alloc(8 * n), ctx.alloc(8 * n), ctx.alloc(8 * n), ctx.alloc(8 * nb * cols)
    in_p, out_p, cols_a = (C.c_void_p * 1)(d_src.p), (C.c_void_p * 1)(out.p), (C.c_uint64 * 1)(cols)

    def issue():
        abi.check(lib.aby3cu_trunc_tuple_at(ctx.h, KEY_A, 4, KEY_B, 9, d_iter.p, stride, d, None, negr.p, rt0.p, rt1.p, n))
        abi.check(lib.aby3cu_gather_rows_multi_at(ctx.h, 1, in_p, cols_a, out_p, d_idx.p, nb, d_iter.p))
        abi.check(lib.aby3cu_counter_add(ctx.h, d_iter.p, 1))

    def expect(it):
        e_negr, e0, e1 = ctx.alloc(8 * n), ctx.alloc(8 * n), ctx.alloc(8 * n)
        abi.check(lib.aby3cu_trunc_tuple(ctx.h, KEY_A, 4 + it * stride, KEY_B, 9 + it * stride, d, None, e_negr.p, e0.p, e1.p, n))
        return [ctx.download(b, n) for b in (e_negr, e0, e1)], src[idx[it * nb:(it + 1) * nb].astype(np.int64)]

    issue()                                      # iteration 0, eagerly
    (a, b, c), g = expect(0)
    assert np.array_equal(ctx.download(negr, n), a) and np.array_equal(ctx.download(rt0, n), b) and np.array_equal(ctx.download(rt1, n), c)
    assert np.array_equal(ctx.download(out, (nb, cols)), g)
    abi.check(lib.aby3cu_capture_begin(ctx.h))
    issue()                                      # recorded, not executed
    exec_ = C.c_void_p()
    abi.check(lib.aby3cu_capture_end(ctx.h, C.byref(exec_)))
    assert int(ctx.download(d_iter, 2, U64)[0]) == 1          # the captured increment did not run
    l0 = ctx.launches
    for it in (1, 2, 3, 4):
        abi.check(lib.aby3cu_graph_launch(ctx.h, exec_, 3))
        (a, b, c), g = expect(it)
        assert np.array_equal(ctx.download(negr, n), a) and np.array_equal(ctx.download(rt0, n), b) and np.array_equal(ctx.download(rt1, n), c), it
        assert np.array_equal(ctx.download(out, (nb, cols)), g), it
    assert int(ctx.download(d_iter, 2, U64)[0]) == 5
    assert ctx.launches - l0 >= 12               # 4 replays x 3 kernels are counted
    abi.check(lib.aby3cu_graph_destroy(exec_))


def test_converter_and_checker_kernels(ctx):
    """bits_expand / bitinj_msgs (Sh3Converter::bitInjection) and bin_check_gates (shadow evaluation) at the ABI level"""
    rows, words, bits = 50, 2, 91
    m0, m1 = rnd(92, rows * words).reshape(rows, words), rnd(93, rows * words).reshape(rows, words)
    d0, d1 = ctx.upload(m0), ctx.upload(m1)
    ex = ctx.alloc(8 * rows * bits)
    abi.check(lib.aby3cu_bits_expand(ctx.h, d0.p, rows, words, bits, ex.p))
    want = np.stack([(m0[:, j // 64] >> (j % 64)) & 1 for j in range(bits)], axis=1)
    assert np.array_equal(ctx.download(ex, (rows, bits)), want)
    n = rows * bits
    o0, o1, msgs = ctx.alloc(8 * n), ctx.alloc(8 * n), ctx.alloc(16 * n)
    abi.check(lib.aby3cu_bitinj_msgs(ctx.h, d0.p, d1.p, rows, words, bits, KEY_A, 3, KEY_B, 8, o0.p, o1.p, msgs.p))
    x0, x1 = o.stream_u64(KEY_A, 3, n), o.stream_u64(KEY_B, 8, n)
    assert np.array_equal(ctx.download(o0, n, U64), x0) and np.array_equal(ctx.download(o1, n, U64), x1)
    b = np.stack([((m0 ^ m1)[:, j // 64] >> (j % 64)) & 1 for j in range(bits)], axis=1).reshape(-1).astype(U64)
    base = (U64(0) - x0 - x1)
    mm = ctx.download(msgs, (n, 2), U64)
    assert np.array_equal(mm[:, 0], base + b) and np.array_equal(mm[:, 1], base + (U64(1) - b))
    # shadow check: planes of a tiny wire memory whose XOR satisfies out = a & b, then one flipped bit
    width = 200
    rb = lib.aby3cu_bin_row_bytes(width)
    rw = rb // 8
    rng = np.random.default_rng(94)
    planes = rng.integers(0, 2**63, (3, 3, rw), dtype=np.int64)           # [plane][wire 0..2][word]
    a = planes[0, 0] ^ planes[1, 0] ^ planes[2, 0]
    bb = planes[0, 1] ^ planes[1, 1] ^ planes[2, 1]
    planes[2, 2] = (a & bb) ^ planes[0, 2] ^ planes[1, 2]
    gates = ctx.upload(np.array([[0, 1, 2, 8]], dtype=np.uint32))
    res = ctx.alloc(16)
    dp = [ctx.upload(planes[k]) for k in range(3)]
    abi.check(lib.aby3cu_bin_check_gates(ctx.h, gates.p, None, 1, dp[0].p, dp[1].p, dp[2].p, rb, width, res.p, res.at(8)))
    assert int(ctx.download(res, 1, U64)[0]) == 0
    planes[1, 2, 0] ^= 1 << 5
    dp[1] = ctx.upload(planes[1])
    abi.check(lib.aby3cu_bin_check_gates(ctx.h, gates.p, None, 1, dp[0].p, dp[1].p, dp[2].p, rb, width, res.p, res.at(8)))
    assert int(ctx.download(res, 1, U64)[0]) == 1
    assert int(ctx.download(res, 4, np.uint32)[2]) == 0                     # first failing gate index


def test_batched_entry_points_equal_the_single_problem_ones(ctx):
    """aby3cu_*_batch: one launch over several problems gives exactly what the single-problem entry points give"""
    def ptrs(bufs):
        return (C.c_void_p * len(bufs))(*[b.p for b in bufs])
    # truncation pairs: 3 jobs with different keys / offsets / shifts / counts
    keys = [bytes(range(k, k + 16)) for k in (1, 20, 40, 60, 80, 100)]
    jobs = [(keys[0], 5, keys[1], 11, 16, 37), (keys[2], 0, keys[3], 3, 23, 128), (keys[4], 1000, keys[5], 7, 16, 1)]
    outs = [[ctx.alloc(8 * j[5]) for _ in range(3)] for j in jobs]
    kn = (C.c_char_p * 3)(*[j[0] for j in jobs]); kp = (C.c_char_p * 3)(*[j[2] for j in jobs])
    en = (C.c_uint64 * 3)(*[j[1] for j in jobs]); ep = (C.c_uint64 * 3)(*[j[3] for j in jobs])
    sh = (C.c_uint64 * 3)(*[j[4] for j in jobs]); cnt = (C.c_uint64 * 3)(*[j[5] for j in jobs])
    abi.check(lib.aby3cu_trunc_tuple_batch_at(ctx.h, 3, kn, en, kp, ep, sh, cnt, ptrs([o_[0] for o_ in outs]), ptrs([o_[1] for o_ in outs]),
                                              ptrs([o_[2] for o_ in outs]), None, 0))
    for j, o_ in zip(jobs, outs):
        e = [ctx.alloc(8 * j[5]) for _ in range(3)]
        abi.check(lib.aby3cu_trunc_tuple(ctx.h, j[0], j[1], j[2], j[3], j[4], None, e[0].p, e[1].p, e[2].p, j[5]))
        for got, want in zip(o_, e):
            assert np.array_equal(ctx.download(got, j[5]), ctx.download(want, j[5]))
    # share ops, transposes, open-and-truncate, GEMV cross term
    n = 1001
    xs, ys = [rnd(200 + k, n) for k in range(6)], [rnd(210 + k, n) for k in range(6)]
    dx, dy, do = [ctx.upload(a) for a in xs], [ctx.upload(a) for a in ys], [ctx.alloc(8 * n) for _ in range(6)]
    for op, f in ((0, lambda a, b: a.view(U64) + b.view(U64)), (1, lambda a, b: a.view(U64) - b.view(U64)), (2, lambda a, b: a.view(U64) ^ b.view(U64))):
        abi.check(lib.aby3cu_share_op_batch(ctx.h, op, 6, ptrs(dx), ptrs(dy), ptrs(do), n))
        for k in range(6):
            assert np.array_equal(ctx.download(do[k], n, U64), f(xs[k], ys[k]))
    r, c = 70, 33
    ms = [rnd(220 + k, r * c).reshape(r, c) for k in range(4)]
    dm, dt = [ctx.upload(m) for m in ms], [ctx.alloc(8 * r * c) for _ in range(4)]
    abi.check(lib.aby3cu_transpose_i64_batch(ctx.h, 4, ptrs(dm), r, c, ptrs(dt)))
    for k in range(4):
        assert np.array_equal(ctx.download(dt[k], (c, r)), ms[k].T)
    s = [[rnd(230 + 4 * k + q, n) for q in range(4)] for k in range(2)]
    ds = [[ctx.upload(a) for a in row] for row in s]
    abi.check(lib.aby3cu_trunc_finish_batch(ctx.h, 2, ptrs([d_[0] for d_ in ds]), ptrs([d_[1] for d_ in ds]), ptrs([d_[2] for d_ in ds]),
                                            ptrs([d_[3] for d_ in ds]), n, 13))
    for k in range(2):
        want = s[k][3].view(U64) + ((s[k][0].view(U64) + s[k][1].view(U64) + s[k][2].view(U64)).view(np.int64) >> 13).view(U64)
        assert np.array_equal(ctx.download(ds[k][3], n, U64), want)
    for (M, K) in ((128, 1024), (1024, 128), (37, 5000), (1, 1)):
        A0 = [rnd(240 + k, M * K).reshape(M, K) for k in range(3)]; A1 = [rnd(250 + k, M * K).reshape(M, K) for k in range(3)]
        B0 = [rnd(260 + k, K).reshape(K, 1) for k in range(3)]; B1 = [rnd(270 + k, K).reshape(K, 1) for k in range(3)]
        C0 = [rnd(280 + k, M).reshape(M, 1) for k in range(3)]
        d = [[ctx.upload(a) for a in grp] for grp in (A0, A1, B0, B1, C0)]
        abi.check(lib.aby3cu_gemv_cross_batch(ctx.h, 3, ptrs(d[0]), ptrs(d[1]), ptrs(d[2]), ptrs(d[3]), M, K, ptrs(d[4]), 1))
        for k in range(3):
            want = o.cross_term(A0[k], A1[k], B0[k], B1[k]).view(U64) + C0[k].view(U64)
            assert np.array_equal(ctx.download(d[4][k], (M, 1), U64), want), (M, K, k)
    # twelve gathers in one launch
    rows, nb = 300, 64
    idx = np.random.default_rng(5).integers(0, rows, nb).astype(np.uint64)
    srcs = [rnd(300 + k, rows * (k % 3 + 1)).reshape(rows, k % 3 + 1) for k in range(12)]
    dsrc, dout = [ctx.upload(a) for a in srcs], [ctx.alloc(8 * nb * (k % 3 + 1)) for k in range(12)]
    cols = (C.c_uint64 * 12)(*[k % 3 + 1 for k in range(12)])
    abi.check(lib.aby3cu_gather_rows_multi(ctx.h, 12, ptrs(dsrc), cols, ptrs(dout), ctx.upload(idx).p, nb))
    for k in range(12):
        assert np.array_equal(ctx.download(dout[k], (nb, k % 3 + 1)), srcs[k][idx.astype(np.int64)])


@pytest.mark.parametrize("gate", [8, 14])
def test_bitwise_rowmajor_equals_the_sliced_definition(ctx, gate):
    """aby3cu_bin_bitwise_rowmajor (one-level bitwise circuit on row-major words, z transposed in registers) against the
    definition: out[j] bit g = f(a, b)[j] bit g ^ (bit j of the zero-share blocks of nonlinear gate and0 + g), keystreams
    from the oracle.  Ragged instance counts, every bit width class, a non-zero first gate index."""
    rng = np.random.default_rng(gate)
    kp, kn = bytes(range(16)), bytes(range(100, 116))
    # (the last two cases are big enough for the wide AES form: four tables + CTR folding over runs of consecutive tiles)
    for n, bits, and0 in ((1, 64, 0), (127, 64, 0), (128, 1, 3), (129, 13, 0), (3000, 32, 1), (2049, 33, 0), (70000, 64, 5), (4096, 63, 0),
                          (600011, 64, 2), (524288, 40, 0)):
        rb = int(lib.aby3cu_bin_row_bytes(n))
        a0, a1, b0, b1 = (rng.integers(-2**63, 2**63, n, dtype=np.int64) for _ in range(4))
        d = [ctx.upload(x) for x in (a0, a1, b0, b1)]
        out, cp = ctx.alloc(8 * n), ctx.alloc(8 * n)
        abi.check(lib.aby3cu_bin_bitwise_rowmajor(ctx.h, gate, d[0].p, d[1].p, d[2].p, d[3].p, out.p, cp.p, n, bits, rb, kp, kn, and0))
        got = ctx.download(out, n)
        assert np.array_equal(ctx.download(cp, n), got)
        z = np.zeros((64, rb), dtype=np.uint8)
        for g in range(bits):
            z[g] = o.keystream(kp, (and0 + g) * rb, rb) ^ o.keystream(kn, (and0 + g) * rb, rb)
        # z is 64 rows (gates) x n columns (instances), LSB first: transpose to n words of 64 bits
        zt = o.bit_transpose(z.reshape(-1), 64, n, rb, 8).view(np.int64)[:n]
        f = (a0 & b0) ^ (a0 & b1) ^ (a1 & b0)
        if gate == 14:
            f ^= a0 ^ b0
        keep = np.int64(-1) if bits == 64 else np.int64((1 << bits) - 1)
        assert np.array_equal(got, (f ^ zt) & keep), (n, bits, and0)


@pytest.mark.parametrize("not_plane", [0, 1, 2])
def test_maxmin_rowmajor_equals_two_bitwise_and_runs(ctx, not_plane):
    """aby3cu_bin_maxmin_rowmajor (aby3-Basic/BoolBasic.cpp:275-312 after the comparison: two int_int_bitwiseAnd(64) runs over
    the stacked 2n-row operands, the NOT between them, the xors of their halves) against the definition, keystreams from the
    oracle: t = f([m; m], [A; B]) ^ z with z bit g of instance j = bit j of key stream block range g of the run's keys;
    min = t1[0:n] ^ t2[n:2n], max = t1[n:2n] ^ t2[0:n].  Ragged n (the two halves of an element sit in different tiles)."""
    rng = np.random.default_rng(40 + not_plane)
    keys = [bytes(range(k, k + 16)) for k in (0, 100, 30, 200)]
    # (the last case is big enough for the four-table AES form of the kernel)
    for n in (1, 63, 64, 100, 1000, 4099, 70001) + ((300007,) if not_plane == 1 else ()):
        rb = int(lib.aby3cu_bin_row_bytes(2 * n))
        c0, c1 = (rng.integers(0, 2, n, dtype=np.int64) for _ in range(2))
        a0, a1, b0, b1 = (rng.integers(-2**63, 2**63, n, dtype=np.int64) for _ in range(4))
        d = [ctx.upload(x) for x in (c0, c1, a0, a1, b0, b1)]
        mn, mx = ctx.alloc(8 * n), ctx.alloc(8 * n)
        abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, keys[0], 0, mn.p, 8 * n))          # the entry clears its outputs itself
        abi.check(lib.aby3cu_bin_maxmin_rowmajor(ctx.h, *[x.p for x in d], mn.p, mx.p, n, rb, keys[0], keys[1], keys[2], keys[3], not_plane))
        m0, m1 = -c0, -c1
        y0, y1 = np.concatenate([a0, b0]), np.concatenate([a1, b1])

        def run(x0, x1, kp, kn):
            z = np.zeros((64, rb), dtype=np.uint8)
            for g in range(64):
                z[g] = o.keystream(kp, g * rb, rb) ^ o.keystream(kn, g * rb, rb)
            zt = o.bit_transpose(z.reshape(-1), 64, 2 * n, rb, 8).view(np.int64)[:2 * n]
            x0, x1 = np.concatenate([x0, x0]), np.concatenate([x1, x1])
            return (x0 & y0) ^ (x0 & y1) ^ (x1 & y0) ^ zt

        t1 = run(m0, m1, keys[0], keys[1])
        t2 = run(~m0 if not_plane == 1 else m0, ~m1 if not_plane == 2 else m1, keys[2], keys[3])
        assert np.array_equal(ctx.download(mn, n), t1[:n] ^ t2[n:]), n
        assert np.array_equal(ctx.download(mx, n), t1[n:] ^ t2[:n]), n


def test_cmpx_gather_scatter_equal_the_indexed_passes(ctx):
    """aby3cu_cmpx_gather / _scatter (one compare-exchange stage of aby3-Basic/Sort.cpp:366-393 on both planes, the range read /
    written once) against the index-vector definition x[i] = src[r0 + 2i], y[i] = src[r0 + d + 2i]."""
    rng = np.random.default_rng(77)
    for L, r0, d in ((2, 0, 1), (148, 0, 1), (148, 1, 127), (148, 1, 1), (5000, 1, 4095), (5000, 1, 3), (70001, 0, 1), (70001, 1, 65535)):
        total = 2 * L
        m = (total - d - r0 + 1) // 2 if total > r0 + d else 0
        src = [rng.integers(-2**63, 2**63, total, dtype=np.int64) for _ in range(2)]
        dsrc = [ctx.upload(a) for a in src]
        outs = [ctx.alloc(8 * max(m, 2)) for _ in range(4)]
        abi.check(lib.aby3cu_cmpx_gather(ctx.h, dsrc[0].p, dsrc[1].p, r0, d, m, outs[0].p, outs[1].p, outs[2].p, outs[3].p))
        ix, iy = r0 + 2 * np.arange(m), r0 + d + 2 * np.arange(m)
        for s in range(2):
            assert np.array_equal(ctx.download(outs[s], m), src[s][ix]), (L, r0, d)
            assert np.array_equal(ctx.download(outs[2 + s], m), src[s][iy]), (L, r0, d)
        nx = [rng.integers(-2**63, 2**63, m, dtype=np.int64) for _ in range(4)]
        dn = [ctx.upload(a) if m else ctx.alloc(16) for a in nx]
        abi.check(lib.aby3cu_cmpx_scatter(ctx.h, dn[0].p, dn[1].p, dn[2].p, dn[3].p, r0, d, m, dsrc[0].p, dsrc[1].p))
        for s in range(2):
            want = src[s].copy()
            want[ix] = nx[s]
            want[iy] = nx[2 + s]
            assert np.array_equal(ctx.download(dsrc[s], total), want), (L, r0, d)


def test_gemv_ring_equals_three_cross_terms(ctx):
    """aby3cu_gemv_ring (the three co-located parties' GEMV cross terms in one launch, every plane of A read once) against the
    oracle's cross term per party on a CONSISTENT sharing (party p's second plane = party p-1's first plane)."""
    ptr3 = lambda bufs: (C.c_void_p * 3)(*[b.p for b in bufs])
    for (M, K) in ((1, 1), (5, 3), (130, 64), (1000, 511), (257, 1024), (300, 2050)):
        A0 = [rnd(500 + p, M * K).reshape(M, K) for p in range(3)]
        B0 = [rnd(510 + p, K).reshape(K, 1) for p in range(3)]
        C0 = [rnd(520 + p, M).reshape(M, 1) for p in range(3)]
        dA, dB0, dC = [ctx.upload(a) for a in A0], [ctx.upload(b) for b in B0], [ctx.upload(c) for c in C0]
        dB1 = [dB0[(p + 2) % 3] for p in range(3)]                 # plane 1 = the previous party's plane 0
        abi.check(lib.aby3cu_gemv_ring(ctx.h, ptr3(dA), ptr3(dB0), ptr3(dB1), M, K, ptr3(dC), 1))
        for p in range(3):
            q = (p + 2) % 3
            want = o.cross_term(A0[p], A0[q], B0[p], B0[q]).view(U64) + C0[p].view(U64)
            assert np.array_equal(ctx.download(dC[p], (M, 1), U64), want), (M, K, p)
        abi.check(lib.aby3cu_gemv_ring(ctx.h, ptr3(dA), ptr3(dB0), ptr3(dB1), M, K, ptr3(dC), 0))
        for p in range(3):
            q = (p + 2) % 3
            assert np.array_equal(ctx.download(dC[p], (M, 1), U64), o.cross_term(A0[p], A0[q], B0[p], B0[q]).view(U64)), (M, K, p)


def test_one_row_to_bit_words(ctx):
    """getOutput of a one-bit bundle (a comparison's result): one wire row, gathered through a row index and optionally
    complemented, becomes one 64-bit word per instance (fast path of aby3cu_bit_transpose_gather) -- against the oracle transpose."""
    rng = np.random.default_rng(9)
    for width in (65, 128, 1000, 70001):
        rb = int(lib.aby3cu_bin_row_bytes(width))
        mem = rng.integers(0, 256, 5 * rb, dtype=np.uint8)
        d_mem = ctx.upload(mem)
        for row, inv in ((0, 0), (3, 0), (4, 1)):
            d_idx, d_inv = ctx.upload(np.array([row], dtype=np.uint32)), ctx.upload(np.array([inv] + [0] * 15, dtype=np.uint8))
            out = ctx.alloc(8 * width)
            abi.check(lib.aby3cu_bit_transpose_gather(ctx.h, d_mem.p, d_idx.p, 1, width, rb, out.p, 8, d_inv.p if inv else None))
            bits = np.unpackbits(mem[row * rb:(row + 1) * rb], bitorder="little")[:width].astype(np.int64)
            assert np.array_equal(ctx.download(out, width), bits ^ inv), (width, row, inv)


def test_batched_linear_level_equals_the_in_order_walk(ctx):
    """aby3cu_bin_linear_plane0 (linear gates of a level in batches of independent gates, operands of a batch loaded before its
    outputs are stored) against aby3cu_bin_level(..., mem1 = NULL) walking the same list in order: random Xor / Nxor / copy
    gates with chains (a gate reading an earlier gate's output), write-after-read pairs, and the greedy batching of
    Sh3BinaryEvaluator::setCir."""
    rng = np.random.default_rng(123)
    for width, wires, n_gates in ((100, 12, 9), (5000, 40, 37), (70001, 64, 120)):
        rb = int(lib.aby3cu_bin_row_bytes(width))
        gates = np.zeros((n_gates, 4), dtype=np.uint32)
        outs_free = list(range(wires // 2, wires))
        rng.shuffle(outs_free)
        for g in range(n_gates):
            t = int(rng.choice([6, 9, 10, 6]))
            o = outs_free[g % len(outs_free)] if g < len(outs_free) else int(rng.integers(wires // 2, wires))
            a, b = int(rng.integers(0, wires)), int(rng.integers(0, wires))
            while a == o:
                a = int(rng.integers(0, wires))
            while b == o or b == a:
                b = int(rng.integers(0, wires))
            gates[g] = (a, b, o, t)
        first = np.ones(n_gates, dtype=np.uint8)
        written, in_batch = [], 0
        for g in range(n_gates):
            a, b, o, _ = gates[g]
            dep = in_batch == 0 or in_batch == 8 or any(w in (a, b, o) for w in written)
            if dep:
                written, in_batch = [], 0
            first[g] = 1 if dep else 0
            written.append(o)
            in_batch += 1
        assert first.sum() < n_gates or n_gates < 3          # (some gates do share a batch)
        mem = rng.integers(0, 256, wires * rb, dtype=np.uint8)
        d_ref, d_new = ctx.upload(mem), ctx.upload(mem)
        dg, df = ctx.upload(gates), ctx.upload(np.concatenate([first, np.zeros(16, dtype=np.uint8)]))
        abi.check(lib.aby3cu_bin_level(ctx.h, dg.p, n_gates, d_ref.p, None, rb, None, None, 0))
        abi.check(lib.aby3cu_bin_linear_plane0(ctx.h, dg.p, n_gates, df.p, d_new.p, rb))
        assert np.array_equal(ctx.download(d_new, wires * rb, np.uint8), ctx.download(d_ref, wires * rb, np.uint8)), (width, n_gates)

