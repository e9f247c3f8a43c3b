// Regression.h -- mini-batch SGD for linear regression over secret shares
// (aby3-ML/Regression.h:14-184).  Same loop as the reference: sample a batch,
// error = XX*w - YY, update = (XX^T * error) >> log2(|B| / lr), w -= update.
// The reference copies the batch row by row on the host (extractBatch, :43-58);
// here the rows are gathered on the device (aby3cu_gather_rows), the transpose and
// the subtractions are device kernels too, so an iteration never touches host data.
#pragma once
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "aby3ML.h"

namespace aby3 {

struct RegressionParam {
    u64 mIterations;
    u64 mBatchSize;
    double mLearningRate;
};

// rows of X (both share planes) selected by `indices` -> XX, on the device
template <Decimal D>
void extractBatch(sf64Matrix<D>& XX, sf64Matrix<D>& YY, const sf64Matrix<D>& X, const sf64Matrix<D>& Y,
                  const u64* d_indices, u64 count) {
    gpu::Context* ctx = gpu::current();
    if (XX.rows() != count || XX.cols() != X.cols()) XX.resize(count, X.cols());
    if (YY.rows() != count || YY.cols() != Y.cols()) YY.resize(count, Y.cols());
    // the same rows of X and Y, both share planes: one launch
    const int64_t* in[4] = {X[0].dev(), X[1].dev(), Y[0].dev(), Y[1].dev()};
    int64_t* out[4] = {XX[0].devOut(), XX[1].devOut(), YY[0].devOut(), YY[1].devOut()};
    const uint64_t cols[4] = {X.cols(), X.cols(), Y.cols(), Y.cols()};
    gpu::check(aby3cu_gather_rows_multi(ctx->h(), 4, in, cols, out, d_indices, count));
}

// batchIndices: mIterations * mBatchSize row indices (the public mini-batch order).  The
// reference derives it from PRNG(toBlock(234543234)) + std::random_shuffle (:24-40,127),
// which is libstdc++-specific; the order is public data, so the caller supplies it.
template <typename Engine, Decimal D>
void SGD_Linear(RegressionParam& params, Engine& engine, sf64Matrix<D>& X, sf64Matrix<D>& Y, sf64Matrix<D>& w,
                const std::vector<u64>& batchIndices) {
    if (X.rows() != Y.rows() || Y.cols() != 1) throw std::runtime_error(LOCATION);
    if (batchIndices.size() != params.mIterations * params.mBatchSize) throw std::runtime_error(LOCATION);
    for (u64 i : batchIndices) if (i >= (u64)X.rows()) throw std::runtime_error("SGD: batch index out of range " LOCATION);
    gpu::Context* ctx = gpu::current();
    gpu::Buffer dIdx(ctx, std::max<size_t>(batchIndices.size() * 8, 16));
    gpu::check(aby3cu_h2d(ctx->h(), dIdx.ptr(), batchIndices.data(), batchIndices.size() * 8));

    sf64Matrix<D> XX(params.mBatchSize, X.cols()), YY(params.mBatchSize, 1);
    // the learning rate in log2 form: this many extra bits are truncated (:139)
    const u64 aB = (u64)std::log2(1 / (params.mLearningRate / params.mBatchSize));

    // ABY3_SGD_TRACE=1: host time spent in each statement of the loop (party threads print at the end)
    const bool trace = std::getenv("ABY3_SGD_TRACE") != nullptr;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto lap = [&](int k, std::chrono::steady_clock::time_point& t) {
        if (!trace) return;
        auto n = now();
        acc[k] += std::chrono::duration<double, std::micro>(n - t).count();
        t = n;
    };
    for (u64 i = 0; i < params.mIterations; ++i) {
        auto t = now();
        extractBatch(XX, YY, X, Y, (const u64*)dIdx.ptr() + i * params.mBatchSize, params.mBatchSize);
        lap(0, t);
        sf64Matrix<D> error = engine.mul(XX, w);             // :157
        lap(1, t);
        error -= YY;
        lap(2, t);
        XX.transposeInPlace();                               // :163
        lap(3, t);
        sf64Matrix<D> update = engine.mulTruncate(XX, error, aB);   // :166
        lap(4, t);
        w = w - update;
        lap(5, t);
    }
    if (trace)
        std::fprintf(stderr, "SGD trace (us/iter, host): extract %.1f  mul %.1f  sub %.1f  transpose %.1f  mulTrunc %.1f  update %.1f\n",
                     acc[0] / params.mIterations, acc[1] / params.mIterations, acc[2] / params.mIterations,
                     acc[3] / params.mIterations, acc[4] / params.mIterations, acc[5] / params.mIterations);
}

// Logistic regression (aby3-ML/Regression.h:218-295): the same loop with the piecewise sigmoid between the two
// products: error = logisticFunc(XX * w) - YY.
template <typename Engine, Decimal D>
void SGD_Logistic(RegressionParam& params, Engine& engine, sf64Matrix<D>& X, sf64Matrix<D>& Y, sf64Matrix<D>& w,
                  const std::vector<u64>& batchIndices) {
    if (X.rows() != Y.rows() || Y.cols() != 1) throw std::runtime_error(LOCATION);
    if (batchIndices.size() != params.mIterations * params.mBatchSize) throw std::runtime_error(LOCATION);
    for (u64 i : batchIndices) if (i >= (u64)X.rows()) throw std::runtime_error("SGD: batch index out of range " LOCATION);
    gpu::Context* ctx = gpu::current();
    gpu::Buffer dIdx(ctx, std::max<size_t>(batchIndices.size() * 8, 16));
    gpu::check(aby3cu_h2d(ctx->h(), dIdx.ptr(), batchIndices.data(), batchIndices.size() * 8));
    sf64Matrix<D> XX(params.mBatchSize, X.cols()), YY(params.mBatchSize, 1);
    const u64 aB = (u64)std::log2(1 / (params.mLearningRate / params.mBatchSize));
    for (u64 i = 0; i < params.mIterations; ++i) {
        extractBatch(XX, YY, X, Y, (const u64*)dIdx.ptr() + i * params.mBatchSize, params.mBatchSize);
        sf64Matrix<D> xw = engine.mul(XX, w);                       // :263
        sf64Matrix<D> fxw = engine.logisticFunc(xw);                // :264
        sf64Matrix<D> error = fxw - YY;                             // :269
        XX.transposeInPlace();                                      // :274
        sf64Matrix<D> update = engine.mulTruncate(XX, error, aB);   // :277
        w = w - update;
    }
}

}  // namespace aby3
