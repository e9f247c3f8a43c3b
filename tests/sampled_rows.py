"""Expected share planes of a 3PC matrix product on SAMPLED output rows, for sizes where the full oracle run takes
minutes (BASELINE configs[1], 4096^3).  The product shards by independent output rows and every keystream is seekable
(CTR), so row i of every party's result needs only: row i of that party's two A planes, its two B planes, and the
keystream words at element offset i * N.  This is a restatement of the same reference lines the C oracle follows
(aby3/sh3/Sh3Evaluator.cpp:92-116, 503-566, 651-730; Sh3ShareGen.h:9-23, 50-75); tests/test_oracle.py pins it against
the C oracle's full-matrix result at small sizes, so it is the oracle evaluated lazily -- not a second opinion."""
import numpy as np

import oracle_lib as o

U64 = np.uint64


def _eval_keys(sess, p):
    """Sh3ShareGen::init (Sh3ShareGen.h:19-20): the AES keys of the two zero-share streams are the first block of the
    prev / next common PRNG."""
    kp = bytes(o.keystream(sess.seed("eval", p, 0), 0, 16))
    kn = bytes(o.keystream(sess.seed("eval", p, 1), 0, 16))
    return kp, kn


def _gather_stream(key, e0, rows, N):
    """u64 words [e0 + r * N, e0 + (r + 1) * N) of keystream `key` for every r in rows"""
    return np.stack([o.stream_u64(key, e0 + int(r) * N, N) for r in rows])


def mul_rows(sess, cursors, A, B, rows):
    """Rows `rows` of orc_mul's result.  A, B: [3][2][..] share arrays; cursors[p] = sess.cursors(p) BEFORE the
    product.  Returns [3][2][len(rows)][N] (plane 1 of party p+1 = plane 0 of party p, Sh3Evaluator.cpp:109-110)."""
    N = B.shape[3]
    out = np.empty((3, 2, len(rows), N), dtype=np.int64)
    for p in range(3):
        kp, kn = _eval_keys(sess, p)
        e0 = int(cursors[p][1])
        z = _gather_stream(kp, e0, rows, N) - _gather_stream(kn, e0, rows, N)          # getShare(), Sh3ShareGen.h:60-75
        c = o.cross_term(A[p, 0][rows], A[p, 1][rows], B[p, 0], B[p, 1]).view(U64)     # :96-99 (matrix form)
        out[p, 0] = (c + z).view(np.int64)
    for p in range(3):
        out[(p + 1) % 3, 1] = out[p, 0]
    return out


def mul_trunc_rows(sess, cursors, A, B, shift, rows):
    """Rows `rows` of orc_mul_trunc's result (randomised truncation pair)."""
    N = B.shape[3]
    out = np.empty((3, 2, len(rows), N), dtype=np.int64)
    vsum = np.zeros((len(rows), N), dtype=U64)
    for p in range(3):
        # getTruncationTuple (:526-537): t0 <- nextCommon, t1 <- prevCommon, arithmetic shifts
        t0 = _gather_stream(sess.seed("eval", p, 1), int(cursors[p][3]) // 8, rows, N).view(np.int64)
        t1 = _gather_stream(sess.seed("eval", p, 0), int(cursors[p][2]) // 8, rows, N).view(np.int64)
        r = t0 >> 2
        out[p, 0] = t0 >> (shift + 2)
        out[p, 1] = t1 >> (shift + 2)
        c = o.cross_term(A[p, 0][rows], A[p, 1][rows], B[p, 0], B[p, 1]).view(U64)     # :662-665
        vsum += c - r.view(U64)                                                        # :672
    s = vsum.view(np.int64) >> shift                                                   # :712-718
    for p in range(2):
        out[p, p] = (out[p, p].view(U64) + s.view(U64)).view(np.int64)
    return out
