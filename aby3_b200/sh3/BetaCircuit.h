// BetaCircuit.h -- the Boolean-circuit container the binary engine consumes, and a
// small library of the circuits the aby3 apps ask for.  In the reference these
// are cryptoTools' oc::BetaCircuit / oc::BetaLibrary (libOTe @ cf537295, not in
// the reference tree; field list per SURVEY 9.2).  The circuit is DATA for the
// engine: gate order and levelisation here are this implementation's own (the
// cryptoTools ones cannot be inspected), reconstructed outputs do not depend on
// them.  Host-only code.
#pragma once
#include <algorithm>
#include <map>

#include "Defines.h"

namespace oc {

// 4-bit truth tables, cryptoTools Gate.h numbering
enum class GateType : u8 {
    Zero = 0, Nor = 1, nb_And = 2, nb = 3, na_And = 4, na = 5, Xor = 6, Nand = 7,
    And = 8, Nxor = 9, a = 10, nb_Or = 11, b = 12, na_Or = 13, Or = 14, One = 15
};

inline bool isLinear(GateType t) { return t == GateType::Xor || t == GateType::Nxor || t == GateType::a; }

typedef u32 BetaWire;

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
    explicit BetaBundle(u64 n) : mWires(n, (BetaWire)-1) {}
    u64 size() const { return mWires.size(); }
    BetaWire& operator[](u64 i) { return mWires[i]; }
    const BetaWire& operator[](u64 i) const { return mWires[i]; }
    BetaWire front() const { return mWires.front(); }
    BetaWire back() const { return mWires.back(); }
};

enum class BetaWireFlag : u8 { Zero, One, Wire, InvWire, Uninitialized };

class BetaCircuit {
public:
    u64 mNonlinearGateCount = 0;
    BetaWire mWireCount = 0;
    std::vector<BetaGate> mGates;
    std::vector<BetaBundle> mInputs, mOutputs;
    std::vector<u64> mLevelCounts, mLevelAndCounts;
    std::vector<BetaWireFlag> mWireFlags;

    void addInputBundle(BetaBundle& in) {
        for (u64 i = 0; i < in.size(); ++i) in[i] = newWire(BetaWireFlag::Wire);
        mInputs.push_back(in);
    }
    void addTempWireBundle(BetaBundle& b) {
        for (u64 i = 0; i < b.size(); ++i) b[i] = newWire(BetaWireFlag::Uninitialized);
    }
    void addOutputBundle(BetaBundle& out) {
        for (u64 i = 0; i < out.size(); ++i) out[i] = newWire(BetaWireFlag::Uninitialized);
        mOutputs.push_back(out);
    }
    // register already-existing wires as an output bundle
    void addOutputWires(const BetaBundle& out) { mOutputs.push_back(out); }
    BetaWire addTempWire() { return newWire(BetaWireFlag::Uninitialized); }

    void addGate(BetaWire in0, BetaWire in1, GateType t, BetaWire out) {
        if (t == GateType::a) { addCopy(in0, out); return; }
        if (in0 >= mWireCount || in1 >= mWireCount || out >= mWireCount) throw RTE_LOC;
        // inverted inputs are folded into the gate type (Xor/Nxor) or materialised
        const bool inv0 = mWireFlags[in0] == BetaWireFlag::InvWire, inv1 = mWireFlags[in1] == BetaWireFlag::InvWire;
        if (inv0 || inv1) throw std::runtime_error("BetaCircuit: gates on inverted wires are not supported; use addInvert on outputs only " LOCATION);
        mGates.emplace_back(in0, in1, t, out);
        mWireFlags[out] = BetaWireFlag::Wire;
        if (!isLinear(t)) ++mNonlinearGateCount;
        mLevelCounts.clear(); mLevelAndCounts.clear();
    }
    void addCopy(BetaWire src, BetaWire dst) {
        mGates.emplace_back(src, src, GateType::a, dst);
        mWireFlags[dst] = BetaWireFlag::Wire;
        mLevelCounts.clear(); mLevelAndCounts.clear();
    }
    // mark a wire as logically inverted; applied when the wire is read as an output
    void addInvert(BetaWire w) {
        if (mWireFlags[w] == BetaWireFlag::Wire) mWireFlags[w] = BetaWireFlag::InvWire;
        else if (mWireFlags[w] == BetaWireFlag::InvWire) mWireFlags[w] = BetaWireFlag::Wire;
        else throw RTE_LOC;
    }
    // dst = NOT src: a copy whose reader must complement it (BetaCircuit::addInvert(in, out))
    void addInvert(BetaWire src, BetaWire dst) {
        addCopy(src, dst);
        mWireFlags[dst] = BetaWireFlag::InvWire;
    }
    bool isInvert(BetaWire w) const { return mWireFlags[w] == BetaWireFlag::InvWire; }

    // Order gates by AND depth.  level(g) = max over inputs of (level of the gate
    // that produced it + 1 if that gate is nonlinear); a nonlinear gate's output
    // is only usable one communication round later (Sh3BinaryEvaluator.cpp:555-573).
    // Inside a level the linear gates come first (construction order), then the nonlinear ones
    // (construction order): a nonlinear gate may read same-level linear outputs but nothing reads a
    // nonlinear output before the next level, so the nonlinear gates of a level are independent.
    void levelByAndDepth() {
        // the level sort below may move a gate past another one: only sound when every wire is written once
        {
            std::vector<u8> written(mWireCount, 0);
            for (auto& in : mInputs) for (auto w : in.mWires) written[w] = 1;
            for (auto& G : mGates) {
                if (written[G.mOutput]) throw std::runtime_error("BetaCircuit::levelByAndDepth: wire " + std::to_string(G.mOutput) +
                                                                 " is written more than once; level the circuit by hand (loadFlat) " LOCATION);
                written[G.mOutput] = 1;
            }
        }
        std::vector<u32> ready(mWireCount, 0);     // first level at which the wire's value is usable
        std::vector<u32> lvl(mGates.size());
        u32 maxLevel = 0;
        for (u64 g = 0; g < mGates.size(); ++g) {
            const auto& G = mGates[g];
            u32 l = std::max(ready[G.mInput[0]], ready[G.mInput[1]]);
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

    // rebuild from the flat description used by the test harness / oracle
    void loadFlat(const u32* gates, u32 gateCount, u32 wireCount, const u32* levelGates, u32 levelCount,
                  const u32* inputFirst, const u32* inputBits, u32 numInputs, const u32* outputOff,
                  const u32* outputBits, const u32* outputWires, const u8* outputInvert, u32 numOutputs) {
        *this = BetaCircuit();
        mWireCount = wireCount;
        mWireFlags.assign(wireCount, BetaWireFlag::Wire);
        for (u32 g = 0; g < gateCount; ++g) {
            mGates.emplace_back(gates[4 * g], gates[4 * g + 1], (GateType)gates[4 * g + 3], gates[4 * g + 2]);
            if (!isLinear(mGates.back().mType)) ++mNonlinearGateCount;
        }
        u64 pos = 0;
        for (u32 l = 0; l < levelCount; ++l) {
            mLevelCounts.push_back(levelGates[l]);
            u64 ands = 0;
            for (u64 g = pos; g < pos + levelGates[l]; ++g) ands += !isLinear(mGates[g].mType);
            mLevelAndCounts.push_back(ands);
            pos += levelGates[l];
        }
        for (u32 k = 0; k < numInputs; ++k) {
            BetaBundle b(inputBits[k]);
            for (u32 i = 0; i < inputBits[k]; ++i) b[i] = inputFirst[k] + i;
            mInputs.push_back(b);
        }
        for (u32 k = 0; k < numOutputs; ++k) {
            BetaBundle b(outputBits[k]);
            for (u32 i = 0; i < outputBits[k]; ++i) {
                b[i] = outputWires[outputOff[k] + i];
                if (outputInvert && outputInvert[outputOff[k] + i]) mWireFlags[b[i]] = BetaWireFlag::InvWire;
            }
            mOutputs.push_back(b);
        }
    }

private:
    BetaWire newWire(BetaWireFlag f) {
        mWireFlags.push_back(f);
        return mWireCount++;
    }
};

// A few of the circuits aby3-Basic / aby3-ML request from oc::BetaLibrary
// (BoolBasic.cpp:29,51,111,152; CircuitLibrary.cpp:379-394).  Built once, cached.
class BetaLibrary {
public:
    enum class Optimized { Size, Depth };
    enum class IntType { TwosComplement, Unsigned };
    enum class AdderType { Addition, Subtraction };
    ~BetaLibrary() { for (auto& kv : mCache) delete kv.second; }

    // c = a + b on the given (already allocated) bundles, c.size() bits, operands sign-extended
    // (BetaLibrary::add_build as called by Sh3Converter.cpp:403-406 and CircuitLibrary.cpp)
    static void add_build(BetaCircuit& cd, const BetaBundle& a, const BetaBundle& b, const BetaBundle& c, const BetaBundle& /*temps*/,
                          IntType it, Optimized op) {
        if (it != IntType::TwosComplement) throw std::runtime_error("add_build: only two's complement operands are used on this path " LOCATION);
        if (op == Optimized::Depth) prefixAdd(cd, a, b, c, false);
        else rippleAdd(cd, a, b, c);
    }

    BetaCircuit* int_int_bitwiseAnd(u64 a, u64 b, u64 c) { return bitwise("and", GateType::And, a, b, c); }
    BetaCircuit* int_int_bitwiseOr(u64 a, u64 b, u64 c) { return bitwise("or", GateType::Or, a, b, c); }
    BetaCircuit* int_int_bitwiseXor(u64 a, u64 b, u64 c) { return bitwise("xor", GateType::Xor, a, b, c); }
    // c_i = NOR(a_i, b_i)  (aby3/Circuit/CircuitLibrary.cpp:397-428, bits_nor_helper)
    BetaCircuit* bits_nor_helper(u64 size) { return bitwise("nor", GateType::Nor, size, size, size); }

    // c = a + b (two's complement, cBits low bits)
    BetaCircuit* int_int_add(u64 aBits, u64 bBits, u64 cBits, Optimized op = Optimized::Size) {
        const std::string key = "add" + std::to_string(aBits) + "_" + std::to_string(bBits) + "_" + std::to_string(cBits) +
                                (op == Optimized::Depth ? "d" : "s");
        return cached(key, [&](BetaCircuit& cd) {
            BetaBundle a(aBits), b(bBits), c(cBits);
            cd.addInputBundle(a); cd.addInputBundle(b); cd.addOutputBundle(c);
            if (op == Optimized::Depth) prefixAdd(cd, a, b, c, false);
            else rippleAdd(cd, a, b, c);
        });
    }
    // c = a - b = a + ~b + 1 (aby3-Basic/BoolBasic.cpp:174): the prefix adder on (a, ~b) with the carry-in folded into bit 0
    BetaCircuit* int_int_subtract(u64 aBits, u64 bBits, u64 cBits, Optimized = Optimized::Size) {
        return cached("sub" + std::to_string(aBits) + "_" + std::to_string(bBits) + "_" + std::to_string(cBits), [&](BetaCircuit& cd) {
            BetaBundle a(aBits), b(bBits), c(cBits);
            cd.addInputBundle(a); cd.addInputBundle(b); cd.addOutputBundle(c);
            prefixAdd(cd, a, b, c, true);
        });
    }
    // the most significant bit of a + b (int_comp_helper / fetch_msb)
    BetaCircuit* int_int_add_msb(u64 bits) {
        return cached("addmsb" + std::to_string(bits), [&](BetaCircuit& cd) {
            BetaBundle a(bits), b(bits), c(1);
            cd.addInputBundle(a); cd.addInputBundle(b); cd.addOutputBundle(c);
            msbOfSum(cd, a, b, c[0]);
        });
    }
    // cryptoTools' BetaLibrary::int_int_lt as this fork's callers rely on it: aby3-Basic feeds (input 0 = B, input 1 = A)
    // for "A < B" (BoolBasic.cpp:29-32) and its test expects bool_cipher_lt(Y, X) to reveal x > y
    // (aby3_tests/BoolTest.cpp:68,122,283) -- the circuit answers "input 1 < input 0".
    BetaCircuit* int_int_lt(u64 aBits, u64 bBits) { return lessThan(aBits, bBits, true); }
    // c = (input 0 < input 1), the order the facade's own callers and tests use
    BetaCircuit* int_int_lt_ab(u64 aBits, u64 bBits) { return lessThan(aBits, bBits, false); }
    // c = (a < b) for signed two's-complement inputs of equal width; swapped: a = input 1, b = input 0
    BetaCircuit* lessThan(u64 aBits, u64 bBits, bool swapped) {
        if (aBits != bBits) throw RTE_LOC;
        return cached((swapped ? "ltswap" : "lt") + std::to_string(aBits), [&](BetaCircuit& cd) {
            const u64 n = aBits;
            BetaBundle in0(n), in1(n), c(1);
            cd.addInputBundle(in0); cd.addInputBundle(in1); cd.addOutputBundle(c);
            const BetaBundle& a = swapped ? in1 : in0;
            const BetaBundle& b = swapped ? in0 : in1;
            // a < b  <=>  sign of (a - b) computed on n+1 bits (sign-extended operands):
            // a - b = a + ~b + 1.  Borrow-chain formulation: lt = MSB(diff_ext).
            // diff_ext bit n = a_s ^ ~b_s ^ carry_n, with carry from a + ~b + 1.
            BetaBundle nb(n);
            cd.addTempWireBundle(nb);
            // carry chain with generate/propagate on (a, ~b), carry-in = 1:
            // g_i = a_i & ~b_i = na_And(b_i, a_i), p_i = a_i ^ ~b_i = Nxor(a_i, b_i)
            std::vector<BetaWire> g(n), p(n);
            for (u64 i = 0; i < n; ++i) {
                g[i] = cd.addTempWire(); p[i] = cd.addTempWire();
                cd.addGate(b[i], a[i], GateType::na_And, g[i]);
                cd.addGate(a[i], b[i], GateType::Nxor, p[i]);
            }
            // carry_n (out of bit n-1) with carry-in 1: fold cin into bit 0: g0' = g0 | p0 = Or(g0,p0)... use
            // g0' = g0 ^ p0 (g0 & p0 == 0 because g = a&~b implies p = a^~b = 0)
            BetaWire g0 = cd.addTempWire();
            cd.addGate(g[0], p[0], GateType::Xor, g0);
            g[0] = g0;
            BetaWire carry = prefixCarry(cd, g, p);
            // sign bit of the (n+1)-bit difference: a_s ^ ~b_s ^ carry = p[n-1] ^ carry
            cd.addGate(p[n - 1], carry, GateType::Xor, c[0]);
        });
    }
    // c = (a == b)
    BetaCircuit* int_eq(u64 bits) {
        return cached("eq" + std::to_string(bits), [&](BetaCircuit& cd) {
            BetaBundle a(bits), b(bits), c(1);
            cd.addInputBundle(a); cd.addInputBundle(b); cd.addOutputBundle(c);
            std::vector<BetaWire> e(bits);
            for (u64 i = 0; i < bits; ++i) { e[i] = cd.addTempWire(); cd.addGate(a[i], b[i], GateType::Nxor, e[i]); }
            while (e.size() > 1) {            // AND tree, depth log2(bits)
                std::vector<BetaWire> nx;
                for (u64 i = 0; i + 1 < e.size(); i += 2) {
                    BetaWire w = cd.addTempWire();
                    cd.addGate(e[i], e[i + 1], GateType::And, w);
                    nx.push_back(w);
                }
                if (e.size() & 1) nx.push_back(e.back());
                e.swap(nx);
            }
            cd.addCopy(e[0], c[0]);
        });
    }

    // aby3/Circuit/CircuitLibrary.cpp:38-137: inputs a_0..a_{T-1} (x0+x1 minus threshold t) and b (x2);
    // region bits: c_0 = [a_0+b < 0], c_t = NOT[a_{t-1}+b < 0] AND [a_t+b < 0], c_T = NOT[a_{T-1}+b < 0].
    BetaCircuit* int_Sh3Piecewise_helper(u64 size, u64 numThresholds) {
        return cached("pw" + std::to_string(size) + "_" + std::to_string(numThresholds), [&](BetaCircuit& cd) {
            std::vector<BetaBundle> aa(numThresholds, BetaBundle(size));
            for (auto& a : aa) cd.addInputBundle(a);
            BetaBundle b(size);
            cd.addInputBundle(b);
            std::vector<BetaBundle> cc(numThresholds + 1, BetaBundle(1));
            for (auto& c : cc) cd.addOutputBundle(c);
            std::vector<BetaWire> th(numThresholds);
            for (u64 t = 0; t < numThresholds; ++t) {
                th[t] = cd.addTempWire();
                msbOfSum(cd, aa[t], b, th[t]);                // sign bit of a_t + b
            }
            cd.addCopy(th[0], cc[0][0]);
            for (u64 t = 1; t < numThresholds; ++t) cd.addGate(th[t - 1], th[t], GateType::na_And, cc[t][0]);
            cd.addInvert(th[numThresholds - 1], cc[numThresholds][0]);
        });
    }
    // aby3/Circuit/CircuitLibrary.cpp:350-394 (int_comp_helper): MSB of a + b
    BetaCircuit* int_comp_helper(u64 size) { return int_int_add_msb(size); }

private:
    std::map<std::string, BetaCircuit*> mCache;

    template <typename F>
    BetaCircuit* cached(const std::string& key, F build) {
        auto it = mCache.find(key);
        if (it != mCache.end()) return it->second;
        auto* cd = new BetaCircuit;
        build(*cd);
        cd->levelByAndDepth();
        mCache[key] = cd;
        return cd;
    }
    BetaCircuit* bitwise(const char* name, GateType t, u64 aBits, u64 bBits, u64 cBits) {
        if (aBits != bBits || aBits != cBits) throw RTE_LOC;
        return cached(std::string(name) + std::to_string(aBits), [&](BetaCircuit& cd) {
            BetaBundle a(aBits), b(bBits), c(cBits);
            cd.addInputBundle(a); cd.addInputBundle(b); cd.addOutputBundle(c);
            for (u64 i = 0; i < aBits; ++i) cd.addGate(a[i], b[i], t, c[i]);
        });
    }
    static BetaWire at(const BetaBundle& x, u64 i) { return x[std::min<u64>(i, x.size() - 1)]; }   // sign extension
    // ripple-carry: 1 AND per bit, depth = bits
    static void rippleAdd(BetaCircuit& cd, const BetaBundle& a, const BetaBundle& b, const BetaBundle& c) {
        const u64 n = c.size();
        BetaWire carry = (BetaWire)-1;
        for (u64 i = 0; i < n; ++i) {
            BetaWire x = at(a, i), y = at(b, i);
            BetaWire axb = cd.addTempWire();
            cd.addGate(x, y, GateType::Xor, axb);
            if (i == 0) {
                cd.addCopy(axb, c[0]);
                if (n > 1) { carry = cd.addTempWire(); cd.addGate(x, y, GateType::And, carry); }
            } else {
                cd.addGate(axb, carry, GateType::Xor, c[i]);
                if (i + 1 < n) {
                    // carry' = carry ^ ((x ^ carry) & (y ^ carry))
                    BetaWire xc = cd.addTempWire(), yc = cd.addTempWire(), t = cd.addTempWire(), nc = cd.addTempWire();
                    cd.addGate(x, carry, GateType::Xor, xc);
                    cd.addGate(y, carry, GateType::Xor, yc);
                    cd.addGate(xc, yc, GateType::And, t);
                    cd.addGate(t, carry, GateType::Xor, nc);
                    carry = nc;
                }
            }
        }
    }
    // Kogge-Stone prefix network on (g, p); returns every prefix generate G[i] = carry out of bit i.
    static std::vector<BetaWire> prefixAll(BetaCircuit& cd, std::vector<BetaWire> g, std::vector<BetaWire> p) {
        const u64 n = g.size();
        for (u64 d = 1; d < n; d <<= 1) {
            std::vector<BetaWire> ng = g, np = p;
            for (u64 i = d; i < n; ++i) {
                // G = g_i ^ (p_i & g_{i-d})   (g_i and p_i & x are never both 1), P = p_i & p_{i-d}
                BetaWire t = cd.addTempWire(), G = cd.addTempWire();
                cd.addGate(p[i], g[i - d], GateType::And, t);
                cd.addGate(g[i], t, GateType::Xor, G);
                ng[i] = G;
                if (i >= 2 * d || true) {
                    BetaWire P = cd.addTempWire();
                    cd.addGate(p[i], p[i - d], GateType::And, P);
                    np[i] = P;
                }
            }
            g.swap(ng); p.swap(np);
        }
        return g;
    }
    // Carry out of the whole (g, p) vector only: a reduction TREE instead of the full Kogge-Stone scan -- the same
    // ceil(log2 n) AND levels, but ~2n instead of ~2n log2 n AND gates.  Comparisons and MSB-of-sum tests (lt,
    // int_comp_helper, the piecewise region tests) need nothing but this one carry.
    // Segment (G, P) of bits [lo, hi): G = carry out given carry-in 0, P = all bits propagate.  Combining a low and a
    // high segment: G = G_hi ^ (P_hi & G_lo)  (never both 1), P = P_hi & P_lo.  A segment that starts at bit 0 is only
    // ever the low operand, so its P is never built.
    static BetaWire prefixCarry(BetaCircuit& cd, const std::vector<BetaWire>& g, const std::vector<BetaWire>& p) {
        struct Seg { BetaWire G, P; bool first; };
        std::vector<Seg> cur(g.size());
        for (u64 i = 0; i < g.size(); ++i) cur[i] = Seg{g[i], p[i], i == 0};
        while (cur.size() > 1) {
            std::vector<Seg> nx;
            for (u64 i = 0; i + 1 < cur.size(); i += 2) {
                const Seg& lo = cur[i];
                const Seg& hi = cur[i + 1];
                Seg c{cd.addTempWire(), (BetaWire)-1, lo.first};
                BetaWire t = cd.addTempWire();
                cd.addGate(hi.P, lo.G, GateType::And, t);
                cd.addGate(hi.G, t, GateType::Xor, c.G);
                if (!lo.first) {
                    c.P = cd.addTempWire();
                    cd.addGate(hi.P, lo.P, GateType::And, c.P);
                }
                nx.push_back(c);
            }
            if (cur.size() & 1) nx.push_back(cur.back());
            cur.swap(nx);
        }
        return cur[0].G;
    }
    // out = most significant bit of a + b (same width): p_msb ^ carry into the top bit, the carry by the reduction tree
    static void msbOfSum(BetaCircuit& cd, const BetaBundle& a, const BetaBundle& b, BetaWire out) {
        const u64 n = a.size();
        if (b.size() != n || n == 0) throw RTE_LOC;
        BetaWire pTop = cd.addTempWire();
        cd.addGate(a[n - 1], b[n - 1], GateType::Xor, pTop);
        if (n == 1) { cd.addCopy(pTop, out); return; }
        std::vector<BetaWire> g(n - 1), p(n - 1);
        for (u64 i = 0; i + 1 < n; ++i) {
            g[i] = cd.addTempWire();
            cd.addGate(a[i], b[i], GateType::And, g[i]);
            if (i) { p[i] = cd.addTempWire(); cd.addGate(a[i], b[i], GateType::Xor, p[i]); }      // p_0 is never used
            else p[i] = (BetaWire)-1;
        }
        BetaWire carry = prefixCarry(cd, g, p);
        cd.addGate(pTop, carry, GateType::Xor, out);
    }
    // depth-optimised adder: log2(bits)+1 AND levels
    // c = a + b, or with subtract: c = a + ~b + 1 (generate a & ~b, propagate a ^ ~b, carry-in folded into bit 0:
    // carry out of bit 0 = g0 | p0 = g0 ^ p0 since g0 & p0 = 0; sum bit 0 = p0 ^ 1)
    static void prefixAdd(BetaCircuit& cd, const BetaBundle& a, const BetaBundle& b, const BetaBundle& c, bool subtract) {
        const u64 n = c.size();
        std::vector<BetaWire> g(n), p(n);
        for (u64 i = 0; i < n; ++i) {
            g[i] = cd.addTempWire(); p[i] = cd.addTempWire();
            if (subtract) {
                cd.addGate(at(b, i), at(a, i), GateType::na_And, g[i]);
                cd.addGate(at(a, i), at(b, i), GateType::Nxor, p[i]);
            } else {
                cd.addGate(at(a, i), at(b, i), GateType::And, g[i]);
                cd.addGate(at(a, i), at(b, i), GateType::Xor, p[i]);
            }
        }
        std::vector<BetaWire> gg = g;
        if (subtract) { gg[0] = cd.addTempWire(); cd.addGate(g[0], p[0], GateType::Xor, gg[0]); }
        std::vector<BetaWire> G = prefixAll(cd, gg, p);
        if (subtract) cd.addInvert(p[0], c[0]);
        else cd.addCopy(p[0], c[0]);
        for (u64 i = 1; i < n; ++i) cd.addGate(p[i], G[i - 1], GateType::Xor, c[i]);
    }
};

}  // namespace oc
