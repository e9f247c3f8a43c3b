// Sh3Encryptor.h -- input sharing and reveal (aby3/sh3/Sh3Encryptor.h:10-170,
// Sh3Encryptor.cpp:15-340, 430-560).  Same entry points, same message pattern
// (share: x0 -> next, recv x1 <- prev; reveal: x0 -> prev, recv <- next); the
// per-element loops run as aby3cu kernels and the messages are device buffers.
// sPackedBin forms: localPackedBinary / remotePackedBinary / reveal.
#pragma once
#include "Sh3FixedPoint.h"
#include "Sh3Runtime.h"
#include "Sh3ShareGen.h"

namespace aby3 {

class Sh3Encryptor {
public:
    void init(u64 partyIdx, block prevSeed, block nextSeed, u64 buffSize = 256) {
        mShareGen.init(prevSeed, nextSeed, buffSize);
        mPartyIdx = partyIdx;
    }
    void init(u64 partyIdx, CommPkg& comm, block seed, u64 buffSize = 256) {
        mShareGen.init(comm, seed, buffSize);
        mPartyIdx = partyIdx;
    }

    // ---- scalars ------------------------------------------------------------
    si64 localInt(CommPkg& comm, i64 val);
    si64 remoteInt(CommPkg& comm);
    Sh3Task localInt(Sh3Task dep, i64 val, si64& dest);
    Sh3Task remoteInt(Sh3Task dep, si64& dest);
    sb64 localBinary(CommPkg& comm, i64 val);
    sb64 remoteBinary(CommPkg& comm);
    Sh3Task localBinary(Sh3Task dep, i64 val, sb64& dest);
    Sh3Task remoteBinary(Sh3Task dep, sb64& dest);
    template <Decimal D>
    Sh3Task localFixed(Sh3Task dep, f64<D> val, sf64<D>& dest) { return localInt(dep, val.mValue, dest.mShare); }
    template <Decimal D>
    Sh3Task remoteFixed(Sh3Task dep, sf64<D>& dest) { return remoteInt(dep, dest.mShare); }

    // ---- matrices -----------------------------------------------------------
    void localIntMatrix(CommPkg& comm, const i64Matrix& m, si64Matrix& dest);
    Sh3Task localIntMatrix(Sh3Task dep, const i64Matrix& m, si64Matrix& dest);
    void remoteIntMatrix(CommPkg& comm, si64Matrix& dest);
    Sh3Task remoteIntMatrix(Sh3Task dep, si64Matrix& dest);
    template <Decimal D>
    Sh3Task localFixedMatrix(Sh3Task dep, const f64Matrix<D>& m, sf64Matrix<D>& dest) {
        return localIntMatrix(dep, m.i64Cast(), dest.i64Cast());
    }
    template <Decimal D>
    Sh3Task remoteFixedMatrix(Sh3Task dep, sf64Matrix<D>& dest) { return remoteIntMatrix(dep, dest.i64Cast()); }
    void localBinMatrix(CommPkg& comm, const i64Matrix& m, sbMatrix& dest);
    Sh3Task localBinMatrix(Sh3Task dep, const i64Matrix& m, sbMatrix& dest);
    void remoteBinMatrix(CommPkg& comm, sbMatrix& dest);
    Sh3Task remoteBinMatrix(Sh3Task dep, sbMatrix& dest);

    // ---- reveal ---------------------------------------------------------------
    i64 reveal(CommPkg& comm, const si64& x);
    i64 revealAll(CommPkg& comm, const si64& x);
    void reveal(CommPkg& comm, u64 partyIdx, const si64& x);
    Sh3Task reveal(Sh3Task dep, const si64& x, i64& dest);
    Sh3Task revealAll(Sh3Task dep, const si64& x, i64& dest);
    Sh3Task reveal(Sh3Task dep, u64 partyIdx, const si64& x);
    i64 reveal(CommPkg& comm, const sb64& x);
    i64 revealAll(CommPkg& comm, const sb64& x);
    void reveal(CommPkg& comm, u64 partyIdx, const sb64& x);
    Sh3Task reveal(Sh3Task dep, const sb64& x, i64& dest);
    Sh3Task revealAll(Sh3Task dep, const sb64& x, i64& dest);
    Sh3Task reveal(Sh3Task dep, u64 partyIdx, const sb64& x);

    void reveal(CommPkg& comm, const si64Matrix& x, i64Matrix& dest);
    void revealAll(CommPkg& comm, const si64Matrix& x, i64Matrix& dest);
    void reveal(CommPkg& comm, u64 partyIdx, const si64Matrix& x);
    i64Matrix revealAll(CommPkg& comm, const si64Matrix& x) { i64Matrix d(x.rows(), x.cols()); revealAll(comm, x, d); return d; }
    void reveal(CommPkg& comm, const sbMatrix& x, i64Matrix& dest);
    void revealAll(CommPkg& comm, const sbMatrix& x, i64Matrix& dest);
    void reveal(CommPkg& comm, u64 partyIdx, const sbMatrix& x);
    i64Matrix revealAll(CommPkg& comm, const sbMatrix& x) { i64Matrix d(x.rows(), x.i64Cols()); revealAll(comm, x, d); return d; }
    Sh3Task reveal(Sh3Task dep, const si64Matrix& x, i64Matrix& dest);
    Sh3Task revealAll(Sh3Task dep, const si64Matrix& x, i64Matrix& dest);
    Sh3Task reveal(Sh3Task dep, u64 partyIdx, const si64Matrix& x);
    Sh3Task reveal(Sh3Task dep, const sbMatrix& x, i64Matrix& dest);
    Sh3Task revealAll(Sh3Task dep, const sbMatrix& x, i64Matrix& dest);
    Sh3Task reveal(Sh3Task dep, u64 partyIdx, const sbMatrix& x);

    template <Decimal D>
    Sh3Task reveal(Sh3Task dep, const sf64<D>& x, f64<D>& dest) { return reveal(dep, x.mShare, dest.mValue); }
    template <Decimal D>
    Sh3Task revealAll(Sh3Task dep, const sf64<D>& x, f64<D>& dest) { return revealAll(dep, x.mShare, dest.mValue); }
    template <Decimal D>
    Sh3Task reveal(Sh3Task dep, u64 partyIdx, const sf64<D>& x) { return reveal(dep, partyIdx, x.mShare); }
    template <Decimal D>
    Sh3Task reveal(Sh3Task dep, const sf64Matrix<D>& x, f64Matrix<D>& dest) { return reveal(dep, x.i64Cast(), dest.i64Cast()); }
    template <Decimal D>
    Sh3Task revealAll(Sh3Task dep, const sf64Matrix<D>& x, f64Matrix<D>& dest) { return revealAll(dep, x.i64Cast(), dest.i64Cast()); }
    template <Decimal D>
    Sh3Task reveal(Sh3Task dep, u64 partyIdx, const sf64Matrix<D>& x) { return reveal(dep, partyIdx, x.i64Cast()); }

    // bit-sliced (sPackedBin) forms -- Sh3Encryptor.cpp:342-425, 627-724: the plaintext rows are transposed
    // into bit-slices before masking; reveal transposes back to one row per secret
    void localPackedBinary(CommPkg& comm, const i64Matrix& m, sPackedBin& dest);
    Sh3Task localPackedBinary(Sh3Task dep, const i64Matrix& m, sPackedBin& dest);
    void remotePackedBinary(CommPkg& comm, sPackedBin& dest);
    Sh3Task remotePackedBinary(Sh3Task dep, sPackedBin& dest);
    void reveal(CommPkg& comm, const sPackedBin& x, i64Matrix& dest);
    void revealAll(CommPkg& comm, const sPackedBin& x, i64Matrix& dest);
    void reveal(CommPkg& comm, u64 partyIdx, const sPackedBin& x);
    Sh3Task reveal(Sh3Task dep, const sPackedBin& x, i64Matrix& dest);
    Sh3Task revealAll(Sh3Task dep, const sPackedBin& x, i64Matrix& dest);
    Sh3Task reveal(Sh3Task dep, u64 partyIdx, const sPackedBin& x);

    void rand(si64Matrix& dest);
    void rand(sbMatrix& dest);
    void rand(sPackedBin& dest);

    u64 mPartyIdx = (u64)-1;
    Sh3ShareGen mShareGen;

private:
    // x0 = m (+|^) z on the device, x0 -> next, post the receive of x1 <- prev
    std::future<void> shareMatrix(CommPkg& comm, const i64Matrix* m, eMatrix<i64>& x0, eMatrix<i64>& x1, bool binary);
    void revealMatrix(CommPkg& comm, const eMatrix<i64>& x0, const eMatrix<i64>& x1, i64Matrix& dest, bool binary);
    std::future<void> sharePacked(CommPkg& comm, const i64Matrix* m, sPackedBin& dest);
    void revealPacked(CommPkg& comm, const sPackedBin& x, i64Matrix& dest);
};

}  // namespace aby3
