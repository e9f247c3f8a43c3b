#pragma once
// shim of cryptoTools/Circuit/BetaCircuit.h: the Boolean-circuit container (fields and builder calls
// the reference touches: Sh3BinaryEvaluator.cpp, aby3/Circuit/CircuitLibrary.cpp).  The circuit is DATA
// for the binary engine; gate order / levelisation here are this shim's own (the originals cannot be
// inspected): linear gates of a level first, then its nonlinear gates, construction order otherwise.
#include <algorithm>
#include <list>
#include "cryptoTools/Common/BitVector.h"
#include "cryptoTools/Common/Defines.h"
namespace osuCrypto {
enum class GateType : u8 {
    Zero = 0, Nor = 1, nb_And = 2, nb = 3, na_And = 4, na = 5, Xor = 6, Nand = 7,
    And = 8, Nxor = 9, a = 10, nb_Or = 11, b = 12, na_Or = 13, Or = 14, One = 15
};
inline bool isLinear(GateType t) { return t == GateType::Xor || t == GateType::Nxor || t == GateType::a ||
                                          t == GateType::Zero || t == GateType::nb || t == GateType::na || t == GateType::b || t == GateType::One; }
inline std::string gateToString(GateType t) {
    static const char* n[16] = {"Zero", "Nor", "nb_And", "nb", "na_And", "na", "Xor", "Nand", "And", "Nxor", "a", "nb_Or", "b", "na_Or", "Or", "One"};
    return n[u8(t) & 15];
}
typedef u32 BetaWire;
enum class BetaWireFlag { Zero, One, Wire, InvWire, Uninitialized };

struct BetaGate {
    std::array<BetaWire, 2> mInput;
    BetaWire mOutput;
    GateType mType;
    BetaGate() = default;
    BetaGate(BetaWire in0, BetaWire in1, GateType t, BetaWire out) : mInput{{in0, in1}}, mOutput(out), mType(t) {}
};
struct BetaBundle {
    std::vector<BetaWire> mWires;
    BetaBundle() = default;
    explicit BetaBundle(u64 n) : mWires(n, BetaWire(-1)) {}
    u64 size() const { return mWires.size(); }
    BetaWire& operator[](u64 i) { return mWires[i]; }
    const BetaWire& operator[](u64 i) const { return mWires[i]; }
    BetaWire& front() { return mWires.front(); }
    const BetaWire& front() const { return mWires.front(); }
    BetaWire& back() { return mWires.back(); }
    const BetaWire& back() const { return mWires.back(); }
};

class BetaCircuit {
public:
    enum class LevelizeType { Reorder, NoReorder };
    struct Print { u64 mGateIdx; BetaWire mWire; std::string mMsg; bool mInvert; };
    typedef std::vector<Print>::iterator PrintIter;

    u64 mNonlinearGateCount = 0;
    BetaWire mWireCount = 0;
    std::vector<BetaGate> mGates;
    std::vector<Print> mPrints;
    std::vector<BetaWireFlag> mWireFlags;
    std::vector<BetaBundle> mInputs, mOutputs;
    std::vector<u64> mLevelCounts, mLevelAndCounts;
    std::string mName;

    void addInputBundle(BetaBundle& in) { for (auto& w : in.mWires) w = newWire(BetaWireFlag::Wire); mInputs.push_back(in); }
    void addTempWireBundle(BetaBundle& t) { for (auto& w : t.mWires) w = newWire(BetaWireFlag::Uninitialized); }
    void addTempWire(BetaWire& w) { w = newWire(BetaWireFlag::Uninitialized); }
    void addOutputBundle(BetaBundle& out) { for (auto& w : out.mWires) w = newWire(BetaWireFlag::Uninitialized); mOutputs.push_back(out); }
    void addConstBundle(BetaBundle& b, const BitVector& val) {
        for (u64 i = 0; i < b.size(); ++i) b[i] = newWire(val[i] ? BetaWireFlag::One : BetaWireFlag::Zero);
    }
    void addConst(BetaWire w, u8 val) { mWireFlags.at(w) = val ? BetaWireFlag::One : BetaWireFlag::Zero; }

    // Inverted / constant inputs are folded away so that every stored gate reads plain wires.
    void addGate(BetaWire in0, BetaWire in1, GateType t, BetaWire out) {
        if (in0 >= mWireCount || in1 >= mWireCount || out >= mWireCount) throw RTE_LOC;
        if (t == GateType::a) { addCopy(in0, out); return; }
        u8 tt = u8(t);
        if (mWireFlags[in0] == BetaWireFlag::Uninitialized || mWireFlags[in1] == BetaWireFlag::Uninitialized) throw RTE_LOC;
        // truth table bit index = (a << 1) | b ... with cryptoTools numbering bit (2*b + a)? -- the
        // encoding used: value at inputs (a, b) is bit (a + 2*b) of the 4-bit type, e.g. And = 0b1000.
        auto flipA = [](u8 x) { return u8(((x & 0b0101) << 1) | ((x & 0b1010) >> 1)); };
        auto flipB = [](u8 x) { return u8(((x & 0b0011) << 2) | ((x & 0b1100) >> 2)); };
        if (mWireFlags[in0] == BetaWireFlag::InvWire) tt = flipA(tt);
        if (mWireFlags[in1] == BetaWireFlag::InvWire) tt = flipB(tt);
        const bool c0 = mWireFlags[in0] == BetaWireFlag::Zero || mWireFlags[in0] == BetaWireFlag::One;
        const bool c1 = mWireFlags[in1] == BetaWireFlag::Zero || mWireFlags[in1] == BetaWireFlag::One;
        if (c0 || c1) throw std::runtime_error("BetaCircuit shim: gates on constant wires are not supported " LOCATION);
        const GateType g = GateType(tt);
        switch (g) {
        case GateType::Xor: case GateType::And: case GateType::Nor: case GateType::Or: case GateType::Nxor: case GateType::na_And:
            mGates.emplace_back(in0, in1, g, out); break;
        case GateType::nb_And:            // a & ~b = na_And with swapped inputs
            mGates.emplace_back(in1, in0, GateType::na_And, out); break;
        default:
            throw std::runtime_error("BetaCircuit shim: gate type not representable for the sh3 engine " LOCATION);
        }
        mWireFlags[out] = BetaWireFlag::Wire;
        if (!isLinear(mGates.back().mType)) ++mNonlinearGateCount;
        mLevelCounts.clear(); mLevelAndCounts.clear();
    }
    void addCopy(BetaWire src, BetaWire dst) {
        const BetaWireFlag f = mWireFlags.at(src);
        if (f == BetaWireFlag::Zero || f == BetaWireFlag::One) { mWireFlags.at(dst) = f; return; }
        mGates.emplace_back(src, src, GateType::a, dst);
        mWireFlags.at(dst) = f;                        // an inverted source stays logically inverted
        mLevelCounts.clear(); mLevelAndCounts.clear();
    }
    void addCopy(const BetaBundle& src, const BetaBundle& dst) { for (u64 i = 0; i < src.size(); ++i) addCopy(src[i], dst[i]); }
    void addInvert(BetaWire w) {
        switch (mWireFlags.at(w)) {
        case BetaWireFlag::Wire: mWireFlags[w] = BetaWireFlag::InvWire; break;
        case BetaWireFlag::InvWire: mWireFlags[w] = BetaWireFlag::Wire; break;
        case BetaWireFlag::Zero: mWireFlags[w] = BetaWireFlag::One; break;
        case BetaWireFlag::One: mWireFlags[w] = BetaWireFlag::Zero; break;
        default: throw RTE_LOC;
        }
    }
    void addInvert(BetaWire src, BetaWire dst) { addCopy(src, dst); addInvert(dst); }
    bool isInvert(BetaWire w) const { return mWireFlags.at(w) == BetaWireFlag::InvWire; }
    bool isConst(BetaWire w) const { return mWireFlags.at(w) == BetaWireFlag::Zero || mWireFlags.at(w) == BetaWireFlag::One; }
    u8 constVal(BetaWire w) const { return mWireFlags.at(w) == BetaWireFlag::One; }

    void addPrint(const BetaBundle& b) { for (u64 i = b.size(); i-- > 0;) addPrint(b[i]); }
    void addPrint(BetaWire w) { mPrints.push_back(Print{u64(mGates.size()), w, "", isInvert(w)}); }
    void addPrint(const std::string& s) { mPrints.push_back(Print{u64(mGates.size()), BetaWire(-1), s, false}); }

    void levelByAndDepth(LevelizeType = LevelizeType::Reorder) {
        std::vector<u32> ready(mWireCount, 0), lvl(mGates.size());
        u32 maxLevel = 0;
        for (u64 g = 0; g < mGates.size(); ++g) {
            const BetaGate& G = mGates[g];
            const u32 l = std::max(ready[G.mInput[0]], ready[G.mInput[1]]);
            lvl[g] = l;
            ready[G.mOutput] = isLinear(G.mType) ? l : l + 1;
            maxLevel = std::max(maxLevel, l);
        }
        std::vector<u64> order(mGates.size());
        for (u64 g = 0; g < order.size(); ++g) order[g] = g;
        std::stable_sort(order.begin(), order.end(), [&](u64 x, u64 y) {
            if (lvl[x] != lvl[y]) return lvl[x] < lvl[y];
            return isLinear(mGates[x].mType) && !isLinear(mGates[y].mType);
        });
        std::vector<BetaGate> sorted(mGates.size());
        mLevelCounts.assign(mGates.empty() ? 0 : maxLevel + 1, 0);
        mLevelAndCounts.assign(mLevelCounts.size(), 0);
        for (u64 k = 0; k < order.size(); ++k) {
            sorted[k] = mGates[order[k]];
            ++mLevelCounts[lvl[order[k]]];
            if (!isLinear(sorted[k].mType)) ++mLevelAndCounts[lvl[order[k]]];
        }
        mGates.swap(sorted);
    }
private:
    BetaWire newWire(BetaWireFlag f) { mWireFlags.push_back(f); return mWireCount++; }
};
}  // namespace osuCrypto
