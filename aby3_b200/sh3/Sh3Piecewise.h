// Sh3Piecewise.h -- piecewise-polynomial evaluation on shares (ReLU, the logistic
// approximation of aby3-ML): aby3/sh3/Sh3Piecewise.{h,cpp}.  The input range is cut
// at public thresholds; a binary circuit turns the secret input into one indicator bit
// per region (MSB-of-sum tests), each region's (degree <= 1, integer slope) polynomial
// is evaluated locally, multiplied by its indicator bit with the bit x arithmetic
// product, and the regions are summed.  All bulk work is on the device.
#pragma once
#include "Sh3BinaryEvaluator.h"
#include "Sh3Evaluator.h"

namespace aby3 {

class Sh3Piecewise {
public:
    struct Coef {
        Coef() = default;
        Coef(const int& i) { *this = i64(i); }
        Coef(const i64& i) { *this = i; }
        Coef(const double& d) { *this = d; }
        bool mIsInteger = false;
        i64 mInt = 0;
        double mDouble = 0;
        void operator=(const i64& i) { mIsInteger = true; mInt = i; }
        void operator=(const int& i) { mIsInteger = true; mInt = i; }
        void operator=(const double& d) { mIsInteger = false; mDouble = d; }
        double getDouble(const u64&) const { return mIsInteger ? static_cast<double>(mInt) : mDouble; }
        i64 getFixedPoint(const u64& dec) const {
            return mIsInteger ? (i64)((u64)mInt * (1ull << dec)) : static_cast<i64>(mDouble * (1ull << dec));
        }
        i64 getInteger() const {
            if (!mIsInteger) throw std::runtime_error(LOCATION);
            return mInt;
        }
    };

    std::vector<Coef> mThresholds;
    std::vector<std::vector<Coef>> mCoefficients;

    // plaintext evaluation on fixed-point integers (Sh3Piecewise.cpp:86-183)
    void eval(const i64Matrix& inputs, i64Matrix& outputs, u64 D, bool print = false);
    template <Decimal D>
    void eval(const f64Matrix<D>& inputs, f64Matrix<D>& outputs, bool print = false) {
        eval(inputs.i64Cast(), outputs.i64Cast(), D, print);
    }
    // region indicators of plaintext inputs: rows x (thresholds + 1), row-major bytes
    std::vector<u8> getInputRegions(const i64Matrix& inputs, u64 D);

    // secret-shared evaluation (Sh3Piecewise.cpp:184-380)
    Sh3Task eval(Sh3Task dep, const si64Matrix& input, si64Matrix& output, u64 D, Sh3Evaluator& evaluator, bool print = false);
    template <Decimal D>
    Sh3Task eval(Sh3Task dep, const sf64Matrix<D>& inputs, sf64Matrix<D>& outputs, Sh3Evaluator& evaluator, bool print = false) {
        return eval(dep, inputs.i64Cast(), outputs.i64Cast(), D, evaluator, print);
    }

    std::vector<sbMatrix> mInputRegions;
    std::vector<sbMatrix> circuitInput0;
    sbMatrix circuitInput1;
    Sh3BinaryEvaluator binEng;
    oc::BetaLibrary lib;
    std::vector<si64Matrix> functionOutputs;

    Sh3Task getInputRegions(const si64Matrix& inputs, u64 decimal, CommPkg& comm, Sh3Task& task, Sh3ShareGen& gen, bool print = false);
    Sh3Task getFunctionValues(const si64Matrix& inputs, CommPkg& comm, Sh3Task self, u64 decimal, span<si64Matrix> outputs);
};

}  // namespace aby3
