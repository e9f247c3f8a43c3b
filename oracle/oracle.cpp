/*
 * oracle/oracle.cpp -- CPU oracle for the ABY3 replicated-share multiplication
 * hot path.  TEST INFRASTRUCTURE ONLY (see oracle.h): never linked or loaded
 * by anything under aby3_b200/.
 *
 * Pinned share-for-share against the reference's own sources compiled from
 * /root/reference (oracle/_ref, tests/test_ref_parity.py).  Still "parity
 * unpinned": the cryptoTools primitives under them (PRNG keystream layout,
 * toBlock byte order, bit transpose; libOTe @
 * cf537295c47a3924c13030a9b796cee9d6ebeace is absent here and restated in
 * oracle/shim).  AES itself is pinned by FIPS-197; reconstruction-level
 * behaviour by re-expressing the reference's unit tests (tests/test_oracle_*.py).
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference).  Written from the algorithm, not copied: the reference
 * uses Eigen / cryptoTools types, this file uses flat int64 arrays.
 */
#include "oracle.h"

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <thread>
#include <vector>

#if defined(__AES__) && defined(__SSE2__)
#include <wmmintrin.h>
#include <emmintrin.h>
#define ORC_AESNI 1
#else
#define ORC_AESNI 0
#endif

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

namespace {

/* ------------------------------------------------------------------------ */
/* AES-128 (FIPS-197).  S-box generated from its definition (inverse in      */
/* GF(2^8) followed by the affine map) so there is no table to mistype.      */
/* ------------------------------------------------------------------------ */
struct AesTables {
    u8 sbox[256];
    AesTables() {
        u8 p = 1, q = 1;
        do {
            /* p *= 3 */
            p = (u8)(p ^ (p << 1) ^ ((p & 0x80) ? 0x1B : 0));
            /* q /= 3 */
            q ^= (u8)(q << 1);
            q ^= (u8)(q << 2);
            q ^= (u8)(q << 4);
            if (q & 0x80) q ^= 0x09;
            u8 x = (u8)(q ^ (u8)((q << 1) | (q >> 7)) ^ (u8)((q << 2) | (q >> 6)) ^
                        (u8)((q << 3) | (q >> 5)) ^ (u8)((q << 4) | (q >> 4)));
            sbox[p] = (u8)(x ^ 0x63);
        } while (p != 1);
        sbox[0] = 0x63;
    }
};
const AesTables& tables() { static AesTables t; return t; }

inline u8 xtime(u8 x) { return (u8)((x << 1) ^ ((x & 0x80) ? 0x1B : 0)); }

struct Aes128 {
    u8 rk[11][16];
#if ORC_AESNI
    __m128i rkv[11];
#endif
    /* oc::AES::setKey: standard key schedule on the 16 key bytes in memory order */
    void setKey(const u8 key[16]) {
        const u8* sb = tables().sbox;
        memcpy(rk[0], key, 16);
        u8 rcon = 1;
        for (int r = 1; r <= 10; ++r) {
            const u8* prev = rk[r - 1];
            u8 t[4] = { sb[prev[13]], sb[prev[14]], sb[prev[15]], sb[prev[12]] };
            t[0] ^= rcon;
            rcon = xtime(rcon);
            for (int c = 0; c < 4; ++c) {
                for (int b = 0; b < 4; ++b) {
                    u8 left = (c == 0) ? t[b] : rk[r][4 * (c - 1) + b];
                    rk[r][4 * c + b] = (u8)(prev[4 * c + b] ^ left);
                }
            }
        }
#if ORC_AESNI
        for (int r = 0; r < 11; ++r) rkv[r] = _mm_loadu_si128((const __m128i*)rk[r]);
#endif
    }
    void encSoft(const u8 in[16], u8 out[16]) const {
        const u8* sb = tables().sbox;
        u8 s[16];
        for (int i = 0; i < 16; ++i) s[i] = (u8)(in[i] ^ rk[0][i]);
        for (int r = 1; r <= 10; ++r) {
            u8 t[16];
            /* SubBytes + ShiftRows: state is column-major, byte i = row i%4, col i/4 */
            for (int c = 0; c < 4; ++c)
                for (int row = 0; row < 4; ++row)
                    t[4 * c + row] = sb[s[4 * ((c + row) & 3) + row]];
            if (r < 10) {
                for (int c = 0; c < 4; ++c) {
                    u8 a0 = t[4 * c], a1 = t[4 * c + 1], a2 = t[4 * c + 2], a3 = t[4 * c + 3];
                    s[4 * c + 0] = (u8)(xtime(a0) ^ (xtime(a1) ^ a1) ^ a2 ^ a3);
                    s[4 * c + 1] = (u8)(a0 ^ xtime(a1) ^ (xtime(a2) ^ a2) ^ a3);
                    s[4 * c + 2] = (u8)(a0 ^ a1 ^ xtime(a2) ^ (xtime(a3) ^ a3));
                    s[4 * c + 3] = (u8)((xtime(a0) ^ a0) ^ a1 ^ a2 ^ xtime(a3));
                }
            } else {
                memcpy(s, t, 16);
            }
            for (int i = 0; i < 16; ++i) s[i] ^= rk[r][i];
        }
        memcpy(out, s, 16);
    }
    inline void enc(const u8 in[16], u8 out[16]) const {
#if ORC_AESNI
        __m128i b = _mm_loadu_si128((const __m128i*)in);
        b = _mm_xor_si128(b, rkv[0]);
        for (int r = 1; r < 10; ++r) b = _mm_aesenc_si128(b, rkv[r]);
        b = _mm_aesenclast_si128(b, rkv[10]);
        _mm_storeu_si128((__m128i*)out, b);
#else
        encSoft(in, out);
#endif
    }
    /* oc::AES::ecbEncCounterMode(baseIdx, n, out): block i = AES(toBlock(baseIdx+i)).
     * toBlock(x) = _mm_set_epi64x(0, x): bytes 0..7 little-endian x, 8..15 zero. */
    void ctr(u64 base, u64 n, u8* out) const {
#if ORC_AESNI
        u64 i = 0;
        for (; i + 4 <= n; i += 4) {
            __m128i b0 = _mm_set_epi64x(0, (long long)(base + i));
            __m128i b1 = _mm_set_epi64x(0, (long long)(base + i + 1));
            __m128i b2 = _mm_set_epi64x(0, (long long)(base + i + 2));
            __m128i b3 = _mm_set_epi64x(0, (long long)(base + i + 3));
            b0 = _mm_xor_si128(b0, rkv[0]); b1 = _mm_xor_si128(b1, rkv[0]);
            b2 = _mm_xor_si128(b2, rkv[0]); b3 = _mm_xor_si128(b3, rkv[0]);
            for (int r = 1; r < 10; ++r) {
                b0 = _mm_aesenc_si128(b0, rkv[r]); b1 = _mm_aesenc_si128(b1, rkv[r]);
                b2 = _mm_aesenc_si128(b2, rkv[r]); b3 = _mm_aesenc_si128(b3, rkv[r]);
            }
            b0 = _mm_aesenclast_si128(b0, rkv[10]); b1 = _mm_aesenclast_si128(b1, rkv[10]);
            b2 = _mm_aesenclast_si128(b2, rkv[10]); b3 = _mm_aesenclast_si128(b3, rkv[10]);
            _mm_storeu_si128((__m128i*)(out + 16 * i), b0);
            _mm_storeu_si128((__m128i*)(out + 16 * (i + 1)), b1);
            _mm_storeu_si128((__m128i*)(out + 16 * (i + 2)), b2);
            _mm_storeu_si128((__m128i*)(out + 16 * (i + 3)), b3);
        }
        for (; i < n; ++i) {
            u8 in[16] = {0};
            u64 c = base + i;
            memcpy(in, &c, 8);
            enc(in, out + 16 * i);
        }
#else
        for (u64 i = 0; i < n; ++i) {
            u8 in[16] = {0};
            u64 c = base + i;
            memcpy(in, &c, 8);
            encSoft(in, out + 16 * i);
        }
#endif
    }
    /* bytes [off, off+n) of AES(0)||AES(1)||... */
    void keystream(u64 off, u64 n, u8* out) const {
        u64 blk = off / 16, skip = off % 16;
        if (skip) {
            u8 tmp[16];
            ctr(blk, 1, tmp);
            u64 take = std::min<u64>(16 - skip, n);
            memcpy(out, tmp + skip, take);
            out += take; n -= take; ++blk;
        }
        u64 full = n / 16;
        ctr(blk, full, out);
        out += 16 * full; n -= 16 * full; blk += full;
        if (n) {
            u8 tmp[16];
            ctr(blk, 1, tmp);
            memcpy(out, tmp, n);
        }
    }
};

/* ------------------------------------------------------------------------ */
/* oc::PRNG restatement (cryptoTools/Crypto/PRNG.h, as used at               */
/* aby3/sh3/Sh3ShareGen.h:11-20 and Sh3Evaluator.cpp:13-14,526-527).         */
/* SetSeed: key = seed, 256-block buffer filled by ecbEncCounterMode from    */
/* counter 0; get() copies bytes out of the buffer; when the buffer runs dry */
/* and >= 8 whole blocks are still wanted they are encrypted straight into   */
/* the destination, then the buffer is refilled.  Net effect: one contiguous */
/* AES-CTR keystream (tests assert this against Aes128::keystream).          */
/* ------------------------------------------------------------------------ */
struct Prng {
    Aes128 aes;
    u8 seed[16];
    std::vector<u8> buf;
    u64 blockIdx = 0, bytesIdx = 0, consumed = 0;
    void SetSeed(const u8 s[16], u64 bufferBlocks = 256) {
        memcpy(seed, s, 16);
        aes.setKey(s);
        blockIdx = 0;
        consumed = 0;
        buf.assign(bufferBlocks * 16, 0);
        refill();
    }
    void refill() {
        aes.ctr(blockIdx, buf.size() / 16, buf.data());
        blockIdx += buf.size() / 16;
        bytesIdx = 0;
    }
    void get(u8* dst, u64 len) {
        consumed += len;
        while (len) {
            u64 step = std::min<u64>(len, buf.size() - bytesIdx);
            memcpy(dst, buf.data() + bytesIdx, step);
            dst += step; len -= step; bytesIdx += step;
            if (bytesIdx == buf.size()) {
                if (len >= 8 * 16) {
                    u64 nb = len / 16;
                    aes.ctr(blockIdx, nb, dst);
                    blockIdx += nb;
                    dst += nb * 16; len -= nb * 16;
                }
                refill();
            }
        }
    }
    void getBlock(u8 out[16]) { get(out, 16); }
};

/* ------------------------------------------------------------------------ */
/* Sh3ShareGen restatement -- aby3/sh3/Sh3ShareGen.h:9-23, 50-92             */
/* ------------------------------------------------------------------------ */
struct ShareGen {
    Prng nextCommon, prevCommon, common;
    Aes128 gen[2];                 /* [0] keyed from prevCommon, [1] from nextCommon */
    std::vector<u8> buff[2];
    u64 shareIdx = 0, shareGenIdx = 0, elems = 0;
    void init(const u8 prevSeed[16], const u8 nextSeed[16], u64 buffSize = 256) {
        /* mCommon.SetSeed(toBlock(3488535245, 2454523)) -- :11 (unused on this path) */
        u8 cs[16];
        u64 lo = 2454523ull, hi = 3488535245ull;
        memcpy(cs, &lo, 8); memcpy(cs + 8, &hi, 8);
        common.SetSeed(cs);
        nextCommon.SetSeed(nextSeed);
        prevCommon.SetSeed(prevSeed);
        shareGenIdx = 0;
        buff[0].assign(buffSize * 16, 0);
        buff[1].assign(buffSize * 16, 0);
        u8 k[16];
        prevCommon.getBlock(k); gen[0].setKey(k);     /* :19 */
        nextCommon.getBlock(k); gen[1].setKey(k);     /* :20 */
        refillBuffer();
        elems = 0;
    }
    void refillBuffer() {                              /* :50-56 */
        gen[0].ctr(shareGenIdx, buff[0].size() / 16, buff[0].data());
        gen[1].ctr(shareGenIdx, buff[1].size() / 16, buff[1].data());
        shareGenIdx += buff[0].size() / 16;
        shareIdx = 0;
    }
    inline void fetch(u64& a, u64& b) {
        if (shareIdx + 8 > buff[0].size()) refillBuffer();
        memcpy(&a, buff[0].data() + shareIdx, 8);
        memcpy(&b, buff[1].data() + shareIdx, 8);
        shareIdx += 8;
        ++elems;
    }
    i64 getShare() { u64 a, b; fetch(a, b); return (i64)(a - b); }        /* :60-75 */
    i64 getBinaryShare() { u64 a, b; fetch(a, b); return (i64)(a ^ b); }  /* :77-92 */
};

/* ------------------------------------------------------------------------ */
/* int64 products over Z_2^64.  i-k-j loop order, blocked over k and j so a  */
/* B panel stays in L2; this is what a scalar/AVX2 Eigen i64 product does.   */
/* ------------------------------------------------------------------------ */
void gemm_rows(const u64* A, const u64* B, u64* C, u64 M, u64 K, u64 N, u64 r0, u64 r1, bool accumulate) {
    (void)M;
    const u64 KB = 256, NB = 1024;
    if (!accumulate)
        for (u64 i = r0; i < r1; ++i) memset(C + i * N, 0, N * 8);
    for (u64 j0 = 0; j0 < N; j0 += NB) {
        u64 j1 = std::min(N, j0 + NB);
        for (u64 k0 = 0; k0 < K; k0 += KB) {
            u64 k1 = std::min(K, k0 + KB);
            for (u64 i = r0; i < r1; ++i) {
                u64* c = C + i * N;
                for (u64 k = k0; k < k1; ++k) {
                    const u64 a = A[i * K + k];
                    const u64* b = B + k * N;
                    for (u64 j = j0; j < j1; ++j) c[j] += a * b[j];
                }
            }
        }
    }
}

void par_rows(u64 M, int workers, const std::function<void(u64, u64)>& f) {
    if (workers <= 1 || M < 2) { f(0, M); return; }
    workers = (int)std::min<u64>((u64)workers, M);
    std::vector<std::thread> th;
    u64 per = (M + workers - 1) / workers;
    for (int w = 0; w < workers; ++w) {
        u64 a = std::min(M, per * w), b = std::min(M, per * (w + 1));
        if (a < b) th.emplace_back([=, &f] { f(a, b); });
    }
    for (auto& t : th) t.join();
}

/* A0*B0 + A0*B1 + A1*B0 as three products (Sh3Evaluator.cpp:96-99, 662-665),
 * or the fork's element-wise form (:101-105, :667-668). */
void cross_term(const u64* A0, const u64* A1, const u64* B0, const u64* B1, u64* C,
                u64 M, u64 K, u64 N, int mode, int workers) {
    if (mode == 1) {
        u64 n = M * N;
        for (u64 i = 0; i < n; ++i) C[i] = A0[i] * B0[i] + A0[i] * B1[i] + A1[i] * B0[i];
        return;
    }
    par_rows(M, workers, [&](u64 r0, u64 r1) {
        gemm_rows(A0, B0, C, M, K, N, r0, r1, false);
        gemm_rows(A0, B1, C, M, K, N, r0, r1, true);
        gemm_rows(A1, B0, C, M, K, N, r0, r1, true);
    });
}

/* SharedOT restatement -- aby3/OT/SharedOT.cpp:6-126.  Sender and helper share (key, idx). */
struct SharedOT {
    Aes128 aes;
    u64 idx = 0;
    void setSeed(const u8 k[16]) { aes.setKey(k); idx = 0; }
    /* send: masked[i][h] = m[i][h] ^ AES(idx+i)[h]   (:6-28) */
    void send(const u64* m, u64* masked, u64 n) {
        std::vector<u64> pad(2 * n);
        aes.ctr(idx, n, (u8*)pad.data());
        idx += n;
        for (u64 i = 0; i < 2 * n; ++i) masked[i] = m[i] ^ pad[i];
    }
    /* help: mc[i] = AES(idx+i)[choice_i]   (:30-94; the 128-block stepping does not change the counters) */
    void help(const u8* choice, u64* mc, u64 n) {
        std::vector<u64> pad(2 * n);
        aes.ctr(idx, n, (u8*)pad.data());
        idx += n;
        for (u64 i = 0; i < n; ++i) mc[i] = pad[2 * i + choice[i]];
    }
    /* recv: out[i] = masked[i][choice_i] ^ mc[i]   (:102-126) */
    static void recv(const u64* masked, const u64* mc, const u8* choice, u64* out, u64 n) {
        for (u64 i = 0; i < n; ++i) out[i] = masked[2 * i + choice[i]] ^ mc[i];
    }
};

struct Party {
    int idx;
    ShareGen enc, eval;
    SharedOT otPrevRecver, otNextRecver;     /* Sh3Evaluator.h:118-124 */
    SharedOT convOT12, convOT02;             /* Sh3Converter.h:19 */
    bool convInit = false;
};

}  // namespace

struct orc_prng { Prng p; };

struct orc_session {
    Party p[3];
    bool disableRandomization = false;
};

namespace {

inline const i64* plane(const i64* s, u64 n, int party, int pl) { return s + ((u64)party * 2 + pl) * n; }
inline i64* plane(i64* s, u64 n, int party, int pl) { return s + ((u64)party * 2 + pl) * n; }

void run_parties(int nthreads, const std::function<void(int)>& f) {
    if (nthreads >= 3) {
        std::thread t0(f, 0), t1(f, 1), t2(f, 2);
        t0.join(); t1.join(); t2.join();
    } else {
        for (int i = 0; i < 3; ++i) f(i);
    }
}

/* Sh3Evaluator::getTruncationTuple -- Sh3Evaluator.cpp:503-566 */
void trunc_tuple(orc_session* s, int party, u64 n, u64 d, i64* R, i64* RT0, i64* RT1) {
    if (s->disableRandomization) {                       /* :506-515 */
        memset(R, 0, n * 8); memset(RT0, 0, n * 8); memset(RT1, 0, n * 8);
        return;
    }
    ShareGen& g = s->p[party].eval;
    g.nextCommon.get((u8*)RT0, n * 8);                   /* :526 */
    g.prevCommon.get((u8*)RT1, n * 8);                   /* :527 */
    const u64 d2 = d + 2;
    for (u64 i = 0; i < n; ++i) {                        /* :528-537, arithmetic shifts */
        R[i] = RT0[i] >> 2;
        RT0[i] >>= d2;
        RT1[i] >>= d2;
    }
}

}  // namespace

extern "C" {

int orc_has_aesni(void) { return ORC_AESNI; }

void orc_aes128_encrypt(const u8 key[16], const u8 in[16], u8 out[16], int force_soft) {
    Aes128 a; a.setKey(key);
    if (force_soft) a.encSoft(in, out); else a.enc(in, out);
}

void orc_aes_ctr_blocks(const u8 key[16], u64 base_idx, u64 nblocks, u8* out) {
    Aes128 a; a.setKey(key); a.ctr(base_idx, nblocks, out);
}

void orc_keystream(const u8 key[16], u64 byte_off, u64 nbytes, u8* out) {
    Aes128 a; a.setKey(key); a.keystream(byte_off, nbytes, out);
}

orc_prng* orc_prng_new(const u8 seed[16], u64 buffer_blocks) {
    orc_prng* p = new orc_prng; p->p.SetSeed(seed, buffer_blocks ? buffer_blocks : 256); return p;
}
void orc_prng_free(orc_prng* p) { delete p; }
void orc_prng_get(orc_prng* p, u8* dst, u64 nbytes) { p->p.get(dst, nbytes); }
u64 orc_prng_bytes_consumed(const orc_prng* p) { return p->p.consumed; }

orc_session* orc_session_new(const u8 enc_seeds[3][2][16], const u8 eval_seeds[3][2][16]) {
    orc_session* s = new orc_session;
    for (int i = 0; i < 3; ++i) {
        s->p[i].idx = i;
        s->p[i].enc.init(enc_seeds[i][0], enc_seeds[i][1]);      /* Sh3Encryptor.h:15 */
        s->p[i].eval.init(eval_seeds[i][0], eval_seeds[i][1]);   /* Sh3Evaluator.cpp:9-15 */
        u8 k[16];
        s->p[i].eval.nextCommon.getBlock(k);   /* mOtPrevRecver.setSeed(mNextCommon.get<block>()) :13 */
        s->p[i].otPrevRecver.setSeed(k);
        s->p[i].eval.prevCommon.getBlock(k);   /* mOtNextRecver.setSeed(mPrevCommon.get<block>()) :14 */
        s->p[i].otNextRecver.setSeed(k);
    }
    return s;
}
void orc_session_free(orc_session* s) { delete s; }
void orc_session_set_disable_randomization(orc_session* s, int on) { s->disableRandomization = on != 0; }

void orc_session_cursors(const orc_session* s, int party, u64 c[6]) {
    const Party& p = s->p[party];
    c[0] = p.enc.elems; c[1] = p.eval.elems;
    c[2] = p.eval.prevCommon.consumed; c[3] = p.eval.nextCommon.consumed;
    c[4] = p.enc.prevCommon.consumed; c[5] = p.enc.nextCommon.consumed;
}

/* Sh3Encryptor::localIntMatrix / remoteIntMatrix -- Sh3Encryptor.cpp:217-279 */
void orc_share_int(orc_session* s, int owner, const i64* plain, i64* sh, u64 n) {
    for (int p = 0; p < 3; ++p) {
        i64* x0 = plane(sh, n, p, 0);
        ShareGen& g = s->p[p].enc;
        for (u64 i = 0; i < n; ++i)
            x0[i] = (i64)((u64)g.getShare() + (p == owner ? (u64)plain[i] : 0ull));
    }
    for (int p = 0; p < 3; ++p)   /* send x0 to next; next stores it as plane 1 */
        memcpy(plane(sh, n, (p + 1) % 3, 1), plane(sh, n, p, 0), n * 8);
}

/* Sh3Encryptor::localBinMatrix / remoteBinMatrix -- Sh3Encryptor.cpp:296-340 */
void orc_share_bin(orc_session* s, int owner, const i64* plain, i64* sh, u64 n) {
    for (int p = 0; p < 3; ++p) {
        i64* x0 = plane(sh, n, p, 0);
        ShareGen& g = s->p[p].enc;
        for (u64 i = 0; i < n; ++i)
            x0[i] = g.getBinaryShare() ^ (p == owner ? plain[i] : 0);
    }
    for (int p = 0; p < 3; ++p)
        memcpy(plane(sh, n, (p + 1) % 3, 1), plane(sh, n, p, 0), n * 8);
}

/* Sh3Encryptor::reveal -- Sh3Encryptor.cpp:497-505 (int), :526-536 (bin):
 * receive next party's plane 0, combine with own two planes. */
void orc_reveal(const i64* sh, u64 n, int party, int binary, i64* out) {
    const i64* a = plane(sh, n, party, 0);
    const i64* b = plane(sh, n, party, 1);
    const i64* c = plane(sh, n, (party + 1) % 3, 0);
    for (u64 i = 0; i < n; ++i)
        out[i] = binary ? (a[i] ^ b[i] ^ c[i]) : (i64)((u64)a[i] + (u64)b[i] + (u64)c[i]);
}

void orc_share_op(const i64* X, const i64* Y, i64* Z, u64 n, int op) {
    u64 tot = 6 * n;
    for (u64 i = 0; i < tot; ++i) {
        u64 x = (u64)X[i], y = (u64)Y[i];
        Z[i] = (i64)(op == 0 ? x + y : op == 1 ? x - y : x ^ y);
    }
}

void orc_plain_mul(const i64* A, const i64* B, i64* C, u64 M, u64 K, u64 N, int mode, int nthreads) {
    if (mode == 1) { for (u64 i = 0; i < M * N; ++i) C[i] = (i64)((u64)A[i] * (u64)B[i]); return; }
    par_rows(M, nthreads, [&](u64 r0, u64 r1) {
        gemm_rows((const u64*)A, (const u64*)B, (u64*)C, M, K, N, r0, r1, false);
    });
}

void orc_cross_term(const i64* A0, const i64* A1, const i64* B0, const i64* B1, i64* C,
                    u64 M, u64 K, u64 N, int mode, int nthreads) {
    cross_term((const u64*)A0, (const u64*)A1, (const u64*)B0, (const u64*)B1, (u64*)C, M, K, N, mode, nthreads);
}

/* Sh3Evaluator::asyncMul(si64Matrix,si64Matrix,si64Matrix) -- Sh3Evaluator.cpp:92-116 */
void orc_mul(orc_session* s, const i64* A, const i64* B, i64* C, u64 M, u64 K, u64 N, int mode, int nthreads) {
    const u64 na = (mode == 1) ? M * N : M * K, nb = (mode == 1) ? M * N : K * N, nc = M * N;
    const int inner = nthreads > 3 ? nthreads / 3 : 1;
    run_parties(nthreads, [&](int p) {
        u64* c0 = (u64*)plane(C, nc, p, 0);
        cross_term((const u64*)plane(A, na, p, 0), (const u64*)plane(A, na, p, 1),
                   (const u64*)plane(B, nb, p, 0), (const u64*)plane(B, nb, p, 1),
                   c0, M, K, N, mode, inner);
        ShareGen& g = s->p[p].eval;
        for (u64 i = 0; i < nc; ++i) c0[i] += (u64)g.getShare();          /* :104 */
    });
    for (int p = 0; p < 3; ++p)                                            /* :109-110 */
        memcpy(plane(C, nc, (p + 1) % 3, 1), plane(C, nc, p, 0), nc * 8);
}

void orc_trunc_tuple(orc_session* s, int party, u64 n, u64 d, i64* R, i64* RT0, i64* RT1) {
    trunc_tuple(s, party, n, d, R, RT0, RT1);
}

/* Sh3Evaluator::asyncMul(si64Matrix,si64Matrix,si64Matrix,shift) -- Sh3Evaluator.cpp:651-730 */
void orc_mul_trunc(orc_session* s, const i64* A, const i64* B, i64* C, u64 M, u64 K, u64 N,
                   int mode, u64 shift, int nthreads) {
    const u64 na = (mode == 1) ? M * N : M * K, nb = (mode == 1) ? M * N : K * N, nc = M * N;
    const int inner = nthreads > 3 ? nthreads / 3 : 1;
    std::vector<u64> v[3];
    run_parties(nthreads, [&](int p) {
        v[p].resize(nc);
        cross_term((const u64*)plane(A, na, p, 0), (const u64*)plane(A, na, p, 1),
                   (const u64*)plane(B, nb, p, 0), (const u64*)plane(B, nb, p, 1),
                   v[p].data(), M, K, N, mode, inner);                     /* :662-668 */
        std::vector<i64> R(nc);
        trunc_tuple(s, p, nc, shift, R.data(), plane(C, nc, p, 0), plane(C, nc, p, 1));  /* :670,673 */
        for (u64 i = 0; i < nc; ++i) v[p][i] -= (u64)R[i];                  /* :672 */
    });
    /* :681-684 open xy-r to parties 0 and 1; :703-718 they add (sum >> shift) to share 0,
     * which is plane 0 at party 0 and plane 1 at party 1 (C.mShares[mPartyIdx]). */
    for (int p = 0; p < 2; ++p) {
        i64* dst = plane(C, nc, p, p);
        for (u64 i = 0; i < nc; ++i) {
            i64 sum = (i64)(v[0][i] + v[1][i] + v[2][i]);
            dst[i] = (i64)((u64)dst[i] + (u64)(sum >> shift));
        }
    }
}

/* Sh3Evaluator::asyncMul(si64Matrix A, sbMatrix B, si64Matrix C) -- Sh3Evaluator.cpp:119-263.
 * A: n x 1 arithmetic shares, B: n x 1 binary shares of ONE bit (bit 0 of each word; the
 * reference asserts the words are 0/1, here bit 0 is taken), C = b * a. */
void orc_mul_bit(orc_session* s, const i64* A, const i64* B, i64* C, u64 n) {
    auto a = [&](int p, int pl) { return (const u64*)plane(A, n, p, pl); };
    auto b = [&](int p, int pl) { return (const u64*)plane(B, n, p, pl); };
    auto c = [&](int p, int pl) { return (u64*)plane(C, n, p, pl); };
    std::vector<u64> s0(2 * n), s1(2 * n), m0(2 * n), m2(2 * n), h0(n), h2(n), r0(n), r1(n);
    std::vector<u8> ch(n), ch2(n);
    /* party 0 (:133-170) */
    {
        ShareGen& g = s->p[0].eval;
        for (u64 i = 0; i < n; ++i) {
            const u64 bb0 = (b(0, 0)[i] ^ b(0, 1)[i]) & 1;
            u64 z, c0, c1;
            g.prevCommon.get((u8*)&z, 8);
            g.nextCommon.get((u8*)&c0, 8);
            g.prevCommon.get((u8*)&c1, 8);
            c(0, 0)[i] = c0; c(0, 1)[i] = c1;
            ch[i] = (u8)(b(0, 0)[i] & 1);
            z = 0 - (c0 + c1) - z;
            s0[2 * i + bb0] = z;
            s0[2 * i + (bb0 ^ 1)] = a(0, 0)[i] + a(0, 1)[i] + z;
        }
        s->p[0].otNextRecver.send(s0.data(), m0.data(), n);       /* sender for the OT to P1 */
        s->p[0].otNextRecver.help(ch.data(), h0.data(), n);       /* helper for P2 -> P1 */
    }
    /* party 2 (:212-258) */
    {
        ShareGen& g = s->p[2].eval;
        for (u64 i = 0; i < n; ++i) {
            const u64 bb1 = (b(2, 0)[i] ^ b(2, 1)[i]) & 1;
            u64 z, c0;
            g.nextCommon.get((u8*)&z, 8);
            g.nextCommon.get((u8*)&c0, 8);
            c(2, 0)[i] = c0;
            ch2[i] = (u8)(b(2, 1)[i] & 1);
            s1[2 * i + bb1] = z;
            s1[2 * i + (bb1 ^ 1)] = a(2, 1)[i] + z;
        }
        s->p[2].otPrevRecver.help(ch2.data(), h2.data(), n);      /* helper for P0 -> P1 */
        s->p[2].otPrevRecver.send(s1.data(), m2.data(), n);       /* sender for the OT to P1 */
    }
    /* party 1 (:171-211) */
    {
        ShareGen& g = s->p[1].eval;
        std::vector<u8> b0(n), b1(n);
        for (u64 i = 0; i < n; ++i) {
            g.prevCommon.get((u8*)&c(1, 1)[i], 8);
            b0[i] = (u8)(b(1, 0)[i] & 1);
            b1[i] = (u8)(b(1, 1)[i] & 1);
        }
        SharedOT::recv(m0.data(), h2.data(), b0.data(), r0.data(), n);   /* b*(a0+a2) - c0 - c2 - z */
        SharedOT::recv(m2.data(), h0.data(), b1.data(), r1.data(), n);   /* b*a1 + z */
        for (u64 i = 0; i < n; ++i) c(1, 0)[i] = r1[i] + r0[i];
        memcpy(c(2, 1), c(1, 0), n * 8);                                  /* send c[0] to P2 */
    }
}

/* Sh3Evaluator::asyncMul(i64 a, sbMatrix B, si64Matrix C) -- Sh3Evaluator.cpp:418-501 */
void orc_mul_bit_pub(orc_session* s, i64 aa, const i64* B, i64* C, u64 n) {
    auto b = [&](int p, int pl) { return (const u64*)plane(B, n, p, pl); };
    auto c = [&](int p, int pl) { return (u64*)plane(C, n, p, pl); };
    std::vector<u64> s0(2 * n), mToP1(2 * n), mToP2(2 * n), h1(n), h2(n);
    std::vector<u8> c1(n), c2(n);
    for (u64 i = 0; i < n; ++i) {                                         /* party 0 :430-441 */
        const u64 bb = (b(0, 0)[i] ^ b(0, 1)[i]) & 1;
        const u64 z = (u64)s->p[0].eval.getShare();
        s0[2 * i + bb] = z;
        s0[2 * i + (bb ^ 1)] = (u64)aa + z;
    }
    s->p[0].otNextRecver.send(s0.data(), mToP1.data(), n);
    s->p[0].otPrevRecver.send(s0.data(), mToP2.data(), n);
    for (u64 i = 0; i < n; ++i) {                                         /* party 1 :455-470 */
        c(1, 1)[i] = (u64)s->p[1].eval.getShare();
        c1[i] = (u8)(b(1, 0)[i] & 1);
    }
    s->p[1].otNextRecver.help(c1.data(), h1.data(), n);                   /* helps P0 -> P2 */
    for (u64 i = 0; i < n; ++i) {                                         /* party 2 :476-491 */
        c(2, 0)[i] = (u64)s->p[2].eval.getShare();
        c2[i] = (u8)(b(2, 1)[i] & 1);
    }
    s->p[2].otPrevRecver.help(c2.data(), h2.data(), n);                   /* helps P0 -> P1 */
    memcpy(c(0, 0), c(1, 1), n * 8);                                      /* P1 -> P0 */
    memcpy(c(0, 1), c(2, 0), n * 8);                                      /* P2 -> P0 */
    SharedOT::recv(mToP1.data(), h2.data(), c1.data(), c(1, 0), n);
    SharedOT::recv(mToP2.data(), h1.data(), c2.data(), c(2, 1), n);
}

/* ------------------------------------------------------------------------ */
/* binary engine                                                             */
/* ------------------------------------------------------------------------ */
/* oc::transpose(MatrixView<u8> in, MatrixView<u8> out): out bit (r,c) = in bit (c,r),
 * LSB-first in each byte (pinned by aby3_tests/Sh3ConverterTests.cpp:12-43). */
void orc_bit_transpose(const u8* in, u64 rows, u64 cols, u64 in_stride, u8* out, u64 out_stride) {
    for (u64 c = 0; c < cols; ++c) {
        u8* o = out + c * out_stride;
        const u64 cb = c >> 3, cm = c & 7;
        for (u64 r = 0; r < rows; ++r) {
            u8 bit = (u8)((in[r * in_stride + cb] >> cm) & 1);
            u8 m = (u8)(1u << (r & 7));
            if (bit) o[r >> 3] |= m; else o[r >> 3] &= (u8)~m;
        }
    }
}

u64 orc_bin_row_bytes(u64 width) { return 256 * ((width + 2047) / 2048); }

void orc_bin_eval(orc_session* s, const orc_circuit* cir, u64 width,
                  const i64* const* inputs, i64* const* outputs, u8* mem_dump) {
    const u64 rb = orc_bin_row_bytes(width);
    const u64 rw = rb / 8;
    const u64 W = cir->wire_count;
    const u64 sendBytes = (width + 7) / 8;
    std::vector<u64> mem[3][2];
    Aes128 aes[3][2];
    u64 shareIdx[3] = {0, 0, 0};
    /* setCir(cir, width, gen): Sh3BinaryEvaluator.h:96-102, .cpp:66-103 */
    for (int p = 0; p < 3; ++p) {
        u8 k[16];
        s->p[p].eval.prevCommon.getBlock(k); aes[p][0].setKey(k);
        s->p[p].eval.nextCommon.getBlock(k); aes[p][1].setKey(k);
        mem[p][0].assign(W * rw, 0);
        mem[p][1].assign(W * rw, 0);
    }
    /* setInput(i, sbMatrix): .cpp:200-253 -- transpose each plane into the wire rows */
    for (u32 k = 0; k < cir->num_inputs; ++k) {
        const u64 bits = cir->input_bits[k], words = (bits + 63) / 64, n = width * words;
        for (int p = 0; p < 3; ++p)
            for (int pl = 0; pl < 2; ++pl) {
                const u8* src = (const u8*)plane(inputs[k], n, p, pl);
                u8* dst = (u8*)(mem[p][pl].data() + (u64)cir->input_first[k] * rw);
                orc_bit_transpose(src, width, bits, words * 8, dst, rb);
            }
    }
    /* roundCallback: .cpp:539-1196 */
    const u32* g = cir->gates;
    std::vector<u8> msg[3];
    std::vector<u32> recvLocs;
    std::vector<u64> z(rw), zt(rw);
    for (u32 lvl = 0; lvl < cir->level_count; ++lvl) {
        const u32 ng = cir->level_gates[lvl];
        recvLocs.clear();
        for (int p = 0; p < 3; ++p) {
            msg[p].clear();
            const u32* gg = g;
            u64* m0 = mem[p][0].data();
            u64* m1 = mem[p][1].data();
            for (u32 j = 0; j < ng; ++j, gg += 4) {
                const u32 in0 = gg[0], in1 = gg[1], out = gg[2], type = gg[3];
                u64* o0 = m0 + (u64)out * rw; u64* o1 = m1 + (u64)out * rw;
                const u64* a0 = m0 + (u64)in0 * rw; const u64* a1 = m1 + (u64)in0 * rw;
                const u64* b0 = m0 + (u64)in1 * rw; const u64* b1 = m1 + (u64)in1 * rw;
                const bool nonlinear = type == ORC_GATE_AND || type == ORC_GATE_OR ||
                                       type == ORC_GATE_NOR || type == ORC_GATE_NA_AND;
                if (nonlinear) {
                    /* getShares(): .cpp:1406-1442 -- 2*simdWidth blocks from each key, xored */
                    aes[p][0].ctr(shareIdx[p], rb / 16, (u8*)zt.data());
                    aes[p][1].ctr(shareIdx[p], rb / 16, (u8*)z.data());
                    for (u64 w = 0; w < rw; ++w) z[w] ^= zt[w];
                    shareIdx[p] += rb / 16;
                    if (p == 0) recvLocs.push_back(out);
                }
                switch (type) {
                case ORC_GATE_XOR:                                           /* :700-728 */
                    for (u64 w = 0; w < rw; ++w) { o0[w] = a0[w] ^ b0[w]; o1[w] = a1[w] ^ b1[w]; }
                    break;
                case ORC_GATE_NXOR:                                          /* :982-1022 */
                    for (u64 w = 0; w < rw; ++w) { o0[w] = ~(a0[w] ^ b0[w]); o1[w] = ~(a1[w] ^ b1[w]); }
                    break;
                case ORC_GATE_COPY:                                          /* :1023-1044 */
                    for (u64 w = 0; w < rw; ++w) { o0[w] = a0[w]; o1[w] = a1[w]; }
                    break;
                case ORC_GATE_AND:                                           /* :729-798 */
                    for (u64 w = 0; w < rw; ++w)
                        o0[w] = (a0[w] & b0[w]) ^ (a0[w] & b1[w]) ^ (a1[w] & b0[w]) ^ z[w];
                    break;
                case ORC_GATE_OR:                                            /* :912-981 */
                    for (u64 w = 0; w < rw; ++w)
                        o0[w] = (a0[w] & b0[w]) ^ (a0[w] & b1[w]) ^ (a1[w] & b0[w]) ^ a0[w] ^ b0[w] ^ z[w];
                    break;
                case ORC_GATE_NOR: {                                         /* :799-911 */
                    for (u64 w = 0; w < rw; ++w) {
                        u64 m00 = ~a0[w], m01 = ~b0[w], m10 = ~a1[w], m11 = ~b1[w];
                        o0[w] = (m11 & m00) ^ (m10 & m01) ^ (m00 & m01) ^ z[w];
                    }
                    break;
                }
                case ORC_GATE_NA_AND:                                        /* :1045-1065 */
                    for (u64 w = 0; w < rw; ++w)
                        o0[w] = (~a0[w] & b0[w]) ^ (~a0[w] & b1[w]) ^ (~a1[w] & b0[w]) ^ z[w];
                    break;
                default:
                    fprintf(stderr, "orc_bin_eval: unsupported gate type %u\n", type);
                    abort();
                }
                if (nonlinear) {
                    size_t off = msg[p].size();
                    msg[p].resize(off + sendBytes);
                    memcpy(msg[p].data() + off, o0, sendBytes);               /* :795-796 */
                }
            }
        }
        g += 4 * (u64)ng;
        /* reshare: send to next, next scatters into plane 1 (:555-573, :1161-1171) */
        for (int p = 0; p < 3; ++p) {
            const int q = (p + 1) % 3;
            const u8* it = msg[p].data();
            for (size_t j = 0; j < recvLocs.size(); ++j, it += sendBytes) {
                u8* row = (u8*)(mem[q][1].data() + (u64)recvLocs[j] * rw);
                memset(row, 0xFF, 32);                                        /* AllOneBlock :567 */
                memcpy(row, it, sendBytes);
            }
        }
    }
    /* getOutput(i, sbMatrix): .cpp:1285-1404 */
    for (u32 k = 0; k < cir->num_outputs; ++k) {
        const u64 bits = cir->output_bits[k], words = (bits + 63) / 64, n = width * words;
        std::vector<u8> tmp(bits * rb);
        for (int p = 0; p < 3; ++p)
            for (int pl = 0; pl < 2; ++pl) {
                for (u64 b = 0; b < bits; ++b) {
                    const u32 wire = cir->output_wires[cir->output_off[k] + b];
                    const u64* src = mem[p][pl].data() + (u64)wire * rw;
                    u64* dst = (u64*)(tmp.data() + b * rb);
                    const bool inv = cir->output_invert && cir->output_invert[cir->output_off[k] + b];
                    for (u64 w = 0; w < rw; ++w) dst[w] = inv ? ~src[w] : src[w];
                }
                u8* o = (u8*)plane(outputs[k], n, p, pl);
                memset(o, 0, n * 8);
                orc_bit_transpose(tmp.data(), bits, width, rb, o, words * 8);
            }
    }
    if (mem_dump) {
        for (int p = 0; p < 3; ++p)
            for (int pl = 0; pl < 2; ++pl)
                memcpy(mem_dump + ((u64)p * 2 + pl) * W * rb, mem[p][pl].data(), W * rb);
    }
}

/* ---- Sh3Converter restatement -- aby3/sh3/Sh3Converter.h:24-41, Sh3Converter.cpp:63-411 -------- */
/* conv.init(rt, eval.mShareGen) on every party: the OT seeds are the next block of the common PRNGs */
void orc_conv_init(orc_session* s) {
    u8 k[16];
    s->p[0].eval.prevCommon.getBlock(k); s->p[0].convOT02.setSeed(k);
    s->p[1].eval.nextCommon.getBlock(k); s->p[1].convOT12.setSeed(k);
    s->p[2].eval.prevCommon.getBlock(k); s->p[2].convOT12.setSeed(k);
    s->p[2].eval.nextCommon.getBlock(k); s->p[2].convOT02.setSeed(k);
    for (auto& p : s->p) p.convInit = true;
}

/* toBinaryMatrix(dep, si64Matrix in, sbMatrix dest) up to the circuit: the two adder inputs.
 * X: arithmetic shares [3][2][n]; x0 / x1: binary shares [3][2][n] of (in.0 + in.2) and in.1
 * (Sh3Converter.cpp:73-196).  dest = adder(x0, x1) is then orc_bin_eval on getArithToBinCircuit. */
void orc_conv_a2b_inputs(orc_session* s, const i64* X, u64 n, i64* x0, i64* x1) {
    memset(x0, 0, 6 * n * 8);
    memset(x1, 0, 6 * n * 8);
    std::vector<i64> r0(n), r2(n);
    s->p[0].eval.prevCommon.get((u8*)r0.data(), n * 8);      /* :90 */
    s->p[2].eval.nextCommon.get((u8*)r2.data(), n * 8);      /* :182 */
    const i64* a0 = plane(X, n, 0, 0);
    const i64* a1 = plane(X, n, 0, 1);
    for (u64 i = 0; i < n; ++i) {
        const i64 v = (i64)((u64)a0[i] + (u64)a1[i]) ^ r0[i];        /* :91 */
        plane(x0, n, 0, 0)[i] = v;
        plane(x0, n, 0, 1)[i] = r0[i];
        plane(x0, n, 1, 1)[i] = v;                                   /* received from party 0, :109/:150 */
        plane(x0, n, 2, 0)[i] = r2[i];
        plane(x1, n, 1, 0)[i] = plane(X, n, 1, 0)[i];                /* :139 */
        plane(x1, n, 2, 1)[i] = plane(X, n, 2, 1)[i];                /* :183 */
    }
}

/* bitInjection(dep, sbMatrix in, si64Matrix dest, twoRounds = false) -- Sh3Converter.cpp:211-370.
 * B: binary shares [3][2][rows * words]; Y: arithmetic shares [3][2][rows * bits], one element per bit. */
void orc_conv_bit_injection(orc_session* s, const i64* B, u64 rows, u64 words, u64 bits, i64* Y) {
    const u64 nin = rows * words, n = rows * bits;
    std::vector<u64> d0(n), d1(n), m(2 * n), masked0(2 * n), masked1(2 * n), mc0(n), mc1(n), out0(n), out1(n);
    std::vector<u8> c0(n), c1(n);
    auto bit = [&](const i64* pl, u64 k) { const u64 i = k / bits, j = k % bits; return (u8)(((u64)pl[i * words + j / 64] >> (j % 64)) & 1); };
    /* party 2, the sender (:318-357) */
    s->p[2].eval.nextCommon.get((u8*)d0.data(), n * 8);
    s->p[2].eval.prevCommon.get((u8*)d1.data(), n * 8);
    for (u64 k = 0; k < n; ++k) {
        const u8 b = bit(plane(B, nin, 2, 0), k) ^ bit(plane(B, nin, 2, 1), k);
        const u64 base = 0 - d0[k] - d1[k];
        m[2 * k] = base; m[2 * k + 1] = base;
        m[2 * k + (b ^ 1)] += 1;
    }
    s->p[2].convOT12.send(m.data(), masked0.data(), n);          /* to receiver 0 */
    s->p[2].convOT02.send(m.data(), masked1.data(), n);          /* to receiver 1 */
    /* party 0: receiver 0 (choices = its own share), helper for receiver 1 (:232-271) */
    for (u64 k = 0; k < n; ++k) c0[k] = bit(plane(B, nin, 0, 0), k);
    s->p[0].convOT02.help(c0.data(), mc1.data(), n);
    s->p[0].eval.prevCommon.get((u8*)plane(Y, n, 0, 1), n * 8);
    /* party 1: helper for receiver 0 with x0 (its prev share), then receiver 1 (:273-314) */
    for (u64 k = 0; k < n; ++k) c1[k] = bit(plane(B, nin, 1, 1), k);
    s->p[1].convOT12.help(c1.data(), mc0.data(), n);
    s->p[1].eval.nextCommon.get((u8*)plane(Y, n, 1, 0), n * 8);
    SharedOT::recv(masked0.data(), mc0.data(), c0.data(), out0.data(), n);
    SharedOT::recv(masked1.data(), mc1.data(), c1.data(), out1.data(), n);
    memcpy(plane(Y, n, 0, 0), out0.data(), n * 8);
    memcpy(plane(Y, n, 1, 1), out1.data(), n * 8);
    memcpy(plane(Y, n, 2, 0), d0.data(), n * 8);
    memcpy(plane(Y, n, 2, 1), d1.data(), n * 8);
}

int orc_selftest(void) {
    /* FIPS-197 Appendix C.1 */
    const u8 key[16] = {0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15};
    const u8 pt[16] = {0x00,0x11,0x22,0x33,0x44,0x55,0x66,0x77,0x88,0x99,0xaa,0xbb,0xcc,0xdd,0xee,0xff};
    const u8 ct[16] = {0x69,0xc4,0xe0,0xd8,0x6a,0x7b,0x04,0x30,0xd8,0xcd,0xb7,0x80,0x70,0xb4,0xc5,0x5a};
    u8 out[16];
    orc_aes128_encrypt(key, pt, out, 1);
    if (memcmp(out, ct, 16)) return 1;
    orc_aes128_encrypt(key, pt, out, 0);
    if (memcmp(out, ct, 16)) return 2;
    /* PRNG == contiguous keystream across refills and direct-encrypt path */
    Prng p; p.SetSeed(key);
    std::vector<u8> a(20000), b(20000);
    p.get(a.data(), 16); p.get(a.data() + 16, 8); p.get(a.data() + 24, 5000); p.get(a.data() + 5024, 14976);
    orc_keystream(key, 0, 20000, b.data());
    if (a != b) return 3;
    return 0;
}

}  // extern "C"
