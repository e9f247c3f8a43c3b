// Sh3Encryptor.cpp -- see Sh3Encryptor.h.  Protocol steps follow
// aby3/sh3/Sh3Encryptor.cpp:21-340 (sharing) and :430-560 (reveal).
#include "Sh3Encryptor.h"

namespace aby3 {

// ------------------------------------------------------------------ scalars ----
si64 Sh3Encryptor::localInt(CommPkg& comm, i64 val) {
    si64 ret;
    ret[0] = (i64)((u64)mShareGen.getShare() + (u64)val);
    comm.mNext.asyncSendCopy(ret[0]);
    comm.mPrev.recv(ret[1]);
    return ret;
}
si64 Sh3Encryptor::remoteInt(CommPkg& comm) { return localInt(comm, 0); }

Sh3Task Sh3Encryptor::localInt(Sh3Task dep, i64 val, si64& dest) {
    return dep.then([this, val, &dest](CommPkg& comm, Sh3Task& self) {
        dest[0] = (i64)((u64)mShareGen.getShare() + (u64)val);
        comm.mNext.asyncSendCopy(dest[0]);
        auto fu = comm.mPrev.asyncRecv(dest[1]);
        self.then([fu = std::move(fu)](CommPkg&, Sh3Task&) mutable { fu.get(); });
    }).getClosure();
}
Sh3Task Sh3Encryptor::remoteInt(Sh3Task dep, si64& dest) { return localInt(dep, 0, dest); }

sb64 Sh3Encryptor::localBinary(CommPkg& comm, i64 val) {
    sb64 ret;
    ret[0] = mShareGen.getBinaryShare() ^ val;
    comm.mNext.asyncSendCopy(ret[0]);
    comm.mPrev.recv(ret[1]);
    return ret;
}
sb64 Sh3Encryptor::remoteBinary(CommPkg& comm) { return localBinary(comm, 0); }

Sh3Task Sh3Encryptor::localBinary(Sh3Task dep, i64 val, sb64& dest) {
    return dep.then([this, val, &dest](CommPkg& comm, Sh3Task& self) {
        dest[0] = mShareGen.getBinaryShare() ^ val;
        comm.mNext.asyncSendCopy(dest[0]);
        auto fu = comm.mPrev.asyncRecv(dest[1]);
        self.then([fu = std::move(fu)](CommPkg&, Sh3Task&) mutable { fu.get(); });
    }).getClosure();
}
Sh3Task Sh3Encryptor::remoteBinary(Sh3Task dep, sb64& dest) { return localBinary(dep, 0, dest); }

// ----------------------------------------------------------------- matrices ----
std::future<void> Sh3Encryptor::shareMatrix(CommPkg& comm, const i64Matrix* m, eMatrix<i64>& x0, eMatrix<i64>& x1, bool binary) {
    gpu::Context* ctx = gpu::current();
    const u64 n = x0.size();
    if (m && m->prefetchPending() && n * sizeof(i64) >= (size_t(1) << 20)) {
        // the plaintext is still on its way (eMatrix::prefetchDevice on a copy stream): the zero share -- keystream only --
        // is drawn first and the plaintext added once it has arrived, instead of the fused kernel waiting for the upload
        mShareGen.getShares(ctx, nullptr, x0.devOut(), n, binary);
        const i64* addend = m->dev();                                 // orders this party's stream behind the upload
        gpu::check(aby3cu_share_op(ctx->h(), binary ? ABY3CU_OP_XOR : ABY3CU_OP_ADD, x0.dev(), addend, x0.devMut(), n));
    } else {
        const i64* addend = m ? m->dev() : nullptr;
        mShareGen.getShares(ctx, addend, x0.devOut(), n, binary);    // Sh3Encryptor.cpp:222-223 / 258-259 / 303-304
    }
    comm.mNext.asyncSendDevice(x0.dev(), n * sizeof(i64));
    x1.resizeLike(x0);
    return comm.mPrev.asyncRecvDevice(x1.devOut(), n * sizeof(i64));
}

void Sh3Encryptor::localIntMatrix(CommPkg& comm, const i64Matrix& m, si64Matrix& ret) {
    if (ret.cols() != m.cols() || ret.size() != m.size()) throw std::runtime_error(LOCATION);
    shareMatrix(comm, &m, ret.mShares[0], ret.mShares[1], false).get();
}
Sh3Task Sh3Encryptor::localIntMatrix(Sh3Task dep, const i64Matrix& m, si64Matrix& ret) {
    return dep.then([this, &m, &ret](CommPkg& comm, Sh3Task& self) {
        if (ret.cols() != m.cols() || ret.size() != m.size()) throw std::runtime_error(LOCATION);
        auto fu = shareMatrix(comm, &m, ret.mShares[0], ret.mShares[1], false);
        self.then([fu = std::move(fu)](CommPkg&, Sh3Task&) mutable { fu.get(); });
    }).getClosure();
}
void Sh3Encryptor::remoteIntMatrix(CommPkg& comm, si64Matrix& ret) {
    shareMatrix(comm, nullptr, ret.mShares[0], ret.mShares[1], false).get();
}
Sh3Task Sh3Encryptor::remoteIntMatrix(Sh3Task dep, si64Matrix& ret) {
    return dep.then([this, &ret](CommPkg& comm, Sh3Task& self) {
        auto fu = shareMatrix(comm, nullptr, ret.mShares[0], ret.mShares[1], false);
        self.then([fu = std::move(fu)](CommPkg&, Sh3Task&) mutable { fu.get(); });
    }).getClosure();
}
void Sh3Encryptor::localBinMatrix(CommPkg& comm, const i64Matrix& m, sbMatrix& ret) {
    if (ret.i64Cols() != m.cols() || ret.i64Size() != m.size()) throw std::runtime_error(LOCATION);
    shareMatrix(comm, &m, ret.mShares[0], ret.mShares[1], true).get();
}
Sh3Task Sh3Encryptor::localBinMatrix(Sh3Task dep, const i64Matrix& m, sbMatrix& ret) {
    return dep.then([this, &m, &ret](CommPkg& comm, Sh3Task& self) {
        if (ret.i64Cols() != m.cols() || ret.i64Size() != m.size()) throw std::runtime_error(LOCATION);
        auto fu = shareMatrix(comm, &m, ret.mShares[0], ret.mShares[1], true);
        self.then([fu = std::move(fu)](CommPkg&, Sh3Task&) mutable { fu.get(); });
    }).getClosure();
}
void Sh3Encryptor::remoteBinMatrix(CommPkg& comm, sbMatrix& ret) {
    shareMatrix(comm, nullptr, ret.mShares[0], ret.mShares[1], true).get();
}
Sh3Task Sh3Encryptor::remoteBinMatrix(Sh3Task dep, sbMatrix& ret) {
    return dep.then([this, &ret](CommPkg& comm, Sh3Task& self) {
        auto fu = shareMatrix(comm, nullptr, ret.mShares[0], ret.mShares[1], true);
        self.then([fu = std::move(fu)](CommPkg&, Sh3Task&) mutable { fu.get(); });
    }).getClosure();
}

// --------------------------------------------------------------- reveal: scalars
i64 Sh3Encryptor::reveal(CommPkg& comm, const si64& x) {
    i64 s;
    comm.mNext.recv(s);
    return (i64)((u64)s + (u64)x[0] + (u64)x[1]);
}
i64 Sh3Encryptor::revealAll(CommPkg& comm, const si64& x) {
    comm.mPrev.asyncSendCopy(x[0]);          // = reveal(comm, prev, x); goes out grouped with the receive below
    return reveal(comm, x);
}
void Sh3Encryptor::reveal(CommPkg& comm, u64 partyIdx, const si64& x) {
    if ((mPartyIdx + 2) % 3 == partyIdx) { comm.mPrev.asyncSendCopy(x[0]); comm.mPrev.flush(); }
}
Sh3Task Sh3Encryptor::reveal(Sh3Task dep, const si64& x, i64& dest) {
    return dep.then([&x, &dest](CommPkg& comm, Sh3Task&) {
        comm.mNext.recv(dest);
        dest = (i64)((u64)dest + (u64)x[0] + (u64)x[1]);
    });
}
Sh3Task Sh3Encryptor::revealAll(Sh3Task dep, const si64& x, i64& dest) {
    reveal(dep, (mPartyIdx + 2) % 3, x);
    return reveal(dep, x, dest);
}
Sh3Task Sh3Encryptor::reveal(Sh3Task dep, u64 partyIdx, const si64& x) {
    const bool send = ((mPartyIdx + 2) % 3) == partyIdx;
    return dep.then([send, &x](CommPkg& comm, Sh3Task&) {
        if (send) comm.mPrev.asyncSendCopy(x[0]);
    });
}
i64 Sh3Encryptor::reveal(CommPkg& comm, const sb64& x) {
    i64 s;
    comm.mNext.recv(s);
    return s ^ x[0] ^ x[1];
}
i64 Sh3Encryptor::revealAll(CommPkg& comm, const sb64& x) {
    comm.mPrev.asyncSendCopy(x[0]);
    return reveal(comm, x);
}
void Sh3Encryptor::reveal(CommPkg& comm, u64 partyIdx, const sb64& x) {
    if ((mPartyIdx + 2) % 3 == partyIdx) { comm.mPrev.asyncSendCopy(x[0]); comm.mPrev.flush(); }
}
Sh3Task Sh3Encryptor::reveal(Sh3Task dep, const sb64& x, i64& dest) {
    return dep.then([&x, &dest](CommPkg& comm, Sh3Task&) {
        comm.mNext.recv(dest);
        dest ^= x[0] ^ x[1];
    });
}
Sh3Task Sh3Encryptor::revealAll(Sh3Task dep, const sb64& x, i64& dest) {
    reveal(dep, (mPartyIdx + 2) % 3, x);
    return reveal(dep, x, dest);
}
Sh3Task Sh3Encryptor::reveal(Sh3Task dep, u64 partyIdx, const sb64& x) {
    const bool send = ((mPartyIdx + 2) % 3) == partyIdx;
    return dep.then([send, &x](CommPkg& comm, Sh3Task&) {
        if (send) comm.mPrev.asyncSendCopy(x[0]);
    });
}

// -------------------------------------------------------------- reveal: matrices
void Sh3Encryptor::revealMatrix(CommPkg& comm, const eMatrix<i64>& x0, const eMatrix<i64>& x1, i64Matrix& dest, bool binary) {
    gpu::Context* ctx = gpu::current();
    const u64 n = x0.size();
    dest.resize(x0.rows(), x0.cols());
    gpu::Buffer tmp(ctx, std::max<size_t>(n * sizeof(i64), 16));
    comm.mNext.recvDevice(tmp.ptr(), n * sizeof(i64));                 // Sh3Encryptor.cpp:502 / :529
    gpu::check(aby3cu_combine3(ctx->h(), binary ? ABY3CU_OP_XOR : ABY3CU_OP_ADD, (const i64*)tmp.ptr(),
                               x0.dev(), x1.dev(), dest.devOut(), n));
}
void Sh3Encryptor::reveal(CommPkg& comm, const si64Matrix& x, i64Matrix& dest) {
    revealMatrix(comm, x.mShares[0], x.mShares[1], dest, false);
}
void Sh3Encryptor::revealAll(CommPkg& comm, const si64Matrix& x, i64Matrix& dest) {
    comm.mPrev.asyncSendDevice(x.mShares[0].dev(), x.size() * sizeof(i64));
    reveal(comm, x, dest);
}
void Sh3Encryptor::reveal(CommPkg& comm, u64 partyIdx, const si64Matrix& x) {
    if ((mPartyIdx + 2) % 3 == partyIdx) {
        comm.mPrev.asyncSendDevice(x.mShares[0].dev(), x.size() * sizeof(i64));
        comm.mPrev.flush();      // a pure send: nothing later in this call would carry it out (NCCL transport)
    }
}
void Sh3Encryptor::reveal(CommPkg& comm, const sbMatrix& x, i64Matrix& dest) {
    revealMatrix(comm, x.mShares[0], x.mShares[1], dest, true);
}
void Sh3Encryptor::revealAll(CommPkg& comm, const sbMatrix& x, i64Matrix& dest) {
    comm.mPrev.asyncSendDevice(x.mShares[0].dev(), x.i64Size() * sizeof(i64));
    reveal(comm, x, dest);
}
void Sh3Encryptor::reveal(CommPkg& comm, u64 partyIdx, const sbMatrix& x) {
    if ((mPartyIdx + 2) % 3 == partyIdx) {
        comm.mPrev.asyncSendDevice(x.mShares[0].dev(), x.i64Size() * sizeof(i64));
        comm.mPrev.flush();
    }
}
Sh3Task Sh3Encryptor::reveal(Sh3Task dep, const si64Matrix& x, i64Matrix& dest) {
    return dep.then([this, &x, &dest](CommPkg& comm, Sh3Task&) { revealMatrix(comm, x.mShares[0], x.mShares[1], dest, false); });
}
Sh3Task Sh3Encryptor::revealAll(Sh3Task dep, const si64Matrix& x, i64Matrix& dest) {
    reveal(dep, (mPartyIdx + 2) % 3, x);
    return reveal(dep, x, dest);
}
Sh3Task Sh3Encryptor::reveal(Sh3Task dep, u64 partyIdx, const si64Matrix& x) {
    const bool send = ((mPartyIdx + 2) % 3) == partyIdx;
    return dep.then([send, &x](CommPkg& comm, Sh3Task&) {
        if (send) comm.mPrev.asyncSendDevice(x.mShares[0].dev(), x.size() * sizeof(i64));
    });
}
Sh3Task Sh3Encryptor::reveal(Sh3Task dep, const sbMatrix& x, i64Matrix& dest) {
    return dep.then([this, &x, &dest](CommPkg& comm, Sh3Task&) { revealMatrix(comm, x.mShares[0], x.mShares[1], dest, true); });
}
Sh3Task Sh3Encryptor::revealAll(Sh3Task dep, const sbMatrix& x, i64Matrix& dest) {
    reveal(dep, (mPartyIdx + 2) % 3, x);
    return reveal(dep, x, dest);
}
Sh3Task Sh3Encryptor::reveal(Sh3Task dep, u64 partyIdx, const sbMatrix& x) {
    const bool send = ((mPartyIdx + 2) % 3) == partyIdx;
    return dep.then([send, &x](CommPkg& comm, Sh3Task&) {
        if (send) comm.mPrev.asyncSendDevice(x.mShares[0].dev(), x.i64Size() * sizeof(i64));
    });
}

// ------------------------------------------------------- bit-sliced sharings ----
// dest.mShares[0] = transpose(m) ^ z (or z alone at the other parties), -> next, <- prev   (Sh3Encryptor.cpp:342-425)
std::future<void> Sh3Encryptor::sharePacked(CommPkg& comm, const i64Matrix* m, sPackedBin& dest) {
    gpu::Context* ctx = gpu::current();
    const u64 n = dest.mShares[0].size();
    if (m) {
        if (dest.bitCount() != m->cols() * 64) throw std::runtime_error(LOCATION);          // :344-347
        if (dest.shareCount() != m->rows()) throw std::runtime_error(LOCATION);
        gpu::Buffer t(ctx, std::max<size_t>(n * 8, 16));
        gpu::check(aby3cu_memset(ctx->h(), t.ptr(), 0, std::max<size_t>(n * 8, 16)));
        if (n) gpu::check(aby3cu_bit_transpose(ctx->h(), m->dev(), m->rows(), dest.bitCount(), m->cols() * 8, t.ptr(),
                                               dest.simdWidth() * 8, nullptr));
        mShareGen.getShares(ctx, (const i64*)t.ptr(), dest.mShares[0].devOut(), n, true);   // :356-357
    } else {
        mShareGen.getShares(ctx, nullptr, dest.mShares[0].devOut(), n, true);               // :403-404
    }
    comm.mNext.asyncSendDevice(dest.mShares[0].dev(), n * sizeof(i64));
    return comm.mPrev.asyncRecvDevice(dest.mShares[1].devOut(), n * sizeof(i64));
}
void Sh3Encryptor::localPackedBinary(CommPkg& comm, const i64Matrix& m, sPackedBin& dest) { sharePacked(comm, &m, dest).get(); }
Sh3Task Sh3Encryptor::localPackedBinary(Sh3Task dep, const i64Matrix& m, sPackedBin& dest) {
    return dep.then([this, &m, &dest](CommPkg& comm, Sh3Task& self) {
        auto fu = sharePacked(comm, &m, dest);
        self.then([fu = std::move(fu)](CommPkg&, Sh3Task&) mutable { fu.get(); });
    }).getClosure();
}
void Sh3Encryptor::remotePackedBinary(CommPkg& comm, sPackedBin& dest) { sharePacked(comm, nullptr, dest).get(); }
Sh3Task Sh3Encryptor::remotePackedBinary(Sh3Task dep, sPackedBin& dest) {
    return dep.then([this, &dest](CommPkg& comm, Sh3Task& self) {
        auto fu = sharePacked(comm, nullptr, dest);
        self.then([fu = std::move(fu)](CommPkg&, Sh3Task&) mutable { fu.get(); });
    }).getClosure();
}
// r = transpose(next.x0 ^ x0 ^ x1): one row per secret, ceil(bitCount / 64) words   (Sh3Encryptor.cpp:627-651, 672-691)
void Sh3Encryptor::revealPacked(CommPkg& comm, const sPackedBin& A, i64Matrix& r) {
    gpu::Context* ctx = gpu::current();
    const u64 n = A.mShares[0].size(), wordWidth = (A.bitCount() + 63) / 64;
    r.resize(A.shareCount(), wordWidth);
    gpu::Buffer buff(ctx, std::max<size_t>(n * 8, 16));
    comm.mNext.recvDevice(buff.ptr(), n * sizeof(i64));
    gpu::check(aby3cu_combine3(ctx->h(), ABY3CU_OP_XOR, (const i64*)buff.ptr(), A.mShares[0].dev(), A.mShares[1].dev(), (i64*)buff.ptr(), n));
    i64* out = r.devOut();
    gpu::check(aby3cu_memset(ctx->h(), out, 0, std::max<size_t>(r.size() * 8, 16)));
    if (n) gpu::check(aby3cu_bit_transpose(ctx->h(), buff.ptr(), A.bitCount(), A.shareCount(), A.simdWidth() * 8, out, wordWidth * 8, nullptr));
}
void Sh3Encryptor::reveal(CommPkg& comm, const sPackedBin& x, i64Matrix& dest) { revealPacked(comm, x, dest); }
void Sh3Encryptor::reveal(CommPkg& comm, u64 partyIdx, const sPackedBin& x) {
    if ((mPartyIdx + 2) % 3 == partyIdx) {
        comm.mPrev.asyncSendDevice(x.mShares[0].dev(), x.mShares[0].size() * sizeof(i64));
        comm.mPrev.flush();
    }
}
void Sh3Encryptor::revealAll(CommPkg& comm, const sPackedBin& x, i64Matrix& dest) {
    comm.mPrev.asyncSendDevice(x.mShares[0].dev(), x.mShares[0].size() * sizeof(i64));
    revealPacked(comm, x, dest);
}
Sh3Task Sh3Encryptor::reveal(Sh3Task dep, const sPackedBin& x, i64Matrix& dest) {
    return dep.then([this, &x, &dest](CommPkg& comm, Sh3Task&) { revealPacked(comm, x, dest); });
}
Sh3Task Sh3Encryptor::reveal(Sh3Task dep, u64 partyIdx, const sPackedBin& x) {
    const bool send = ((mPartyIdx + 2) % 3) == partyIdx;
    return dep.then([send, &x](CommPkg& comm, Sh3Task&) {
        if (send) comm.mPrev.asyncSendDevice(x.mShares[0].dev(), x.mShares[0].size() * sizeof(i64));
    });
}
Sh3Task Sh3Encryptor::revealAll(Sh3Task dep, const sPackedBin& x, i64Matrix& dest) {
    reveal(dep, (mPartyIdx + 2) % 3, x);
    return reveal(dep, x, dest);
}

// Sh3Encryptor::rand (Sh3Encryptor.cpp:726-758): a fresh random sharing, plane 0 from
// the next-key stream and plane 1 from the prev-key stream (getRandIntShare).
void Sh3Encryptor::rand(si64Matrix& dest) {
    gpu::Context* ctx = gpu::current();
    const u64 n = dest.size();
    gpu::check(aby3cu_aes_ctr_fill(ctx->h(), mShareGen.mShareGen[1].key().data(), 8 * mShareGen.mShareElemIdx, dest.mShares[0].devOut(), 8 * n));
    gpu::check(aby3cu_aes_ctr_fill(ctx->h(), mShareGen.mShareGen[0].key().data(), 8 * mShareGen.mShareElemIdx, dest.mShares[1].devOut(), 8 * n));
    mShareGen.mShareElemIdx += n;
}
void Sh3Encryptor::rand(sPackedBin& dest) {
    gpu::Context* ctx = gpu::current();
    const u64 n = dest.mShares[0].size();
    gpu::check(aby3cu_aes_ctr_fill(ctx->h(), mShareGen.mShareGen[1].key().data(), 8 * mShareGen.mShareElemIdx, dest.mShares[0].devOut(), 8 * n));
    gpu::check(aby3cu_aes_ctr_fill(ctx->h(), mShareGen.mShareGen[0].key().data(), 8 * mShareGen.mShareElemIdx, dest.mShares[1].devOut(), 8 * n));
    mShareGen.mShareElemIdx += n;
}
void Sh3Encryptor::rand(sbMatrix& dest) {
    gpu::Context* ctx = gpu::current();
    const u64 n = dest.i64Size();
    gpu::check(aby3cu_aes_ctr_fill(ctx->h(), mShareGen.mShareGen[1].key().data(), 8 * mShareGen.mShareElemIdx, dest.mShares[0].devOut(), 8 * n));
    gpu::check(aby3cu_aes_ctr_fill(ctx->h(), mShareGen.mShareGen[0].key().data(), 8 * mShareGen.mShareElemIdx, dest.mShares[1].devOut(), 8 * n));
    mShareGen.mShareElemIdx += n;
}

}  // namespace aby3
