#pragma once
// shim of libOTe/Tools/Tools.h: oc::transpose(MatrixView<u8> in, MatrixView<u8> out) -- bit-matrix
// transpose, LSB-first inside bytes: out bit (r, c) = in bit (c, r) for r < 8*in.cols(), c < in.rows().
// (The bit order is pinned by the reference itself: aby3_tests/Sh3ConverterTests.cpp:12-43, 167-284.)
#include "cryptoTools/Common/MatrixView.h"
namespace osuCrypto {
inline void transpose(const MatrixView<u8>& in, const MatrixView<u8>& out) {
    const u64 inRows = in.rows(), inBits = in.cols() * 8;
    // the original requires the output to be large enough; bits beyond the input are cleared
    const u64 outRows = std::min<u64>(out.rows(), inBits);
    if (out.cols() * 8 < inRows) throw std::runtime_error("transpose shim: output rows too short " LOCATION);
    for (u64 r = 0; r < out.rows(); ++r) std::memset(out.data() + r * out.stride(), 0, out.cols());
    for (u64 c = 0; c < inRows; ++c) {
        const u8* src = in.data() + c * in.stride();
        for (u64 r = 0; r < outRows; ++r)
            if ((src[r >> 3] >> (r & 7)) & 1) out.data()[r * out.stride() + (c >> 3)] |= u8(1u << (c & 7));
    }
}
}  // namespace osuCrypto
