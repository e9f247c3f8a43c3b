#pragma once
// shim of cryptoTools/Circuit/BetaLibrary.h: the builders the reference calls (CircuitLibrary.cpp,
// Sh3Converter.cpp:374-405, aby3-Basic).  Constructions are this shim's own (ripple adder for
// Optimized::Size, Kogge-Stone prefix adder for Optimized::Depth); reconstructed values are what the
// reference's call sites rely on, gate counts/orders are NOT claimed to equal cryptoTools'.
#include <unordered_map>
#include "cryptoTools/Circuit/BetaCircuit.h"
namespace osuCrypto {
class BetaLibrary {
public:
    enum class Optimized { Size, Depth };
    enum class IntType { TwosComplement, Unsigned };
    enum class AdderType { Addition, Subtraction };
    std::unordered_map<size_t, BetaCircuit*> mCirMap;
    std::unordered_map<std::string, BetaCircuit*> mNamed;
    BetaLibrary() = default;
    BetaLibrary(const BetaLibrary&) = delete;
    ~BetaLibrary() { for (auto& kv : mCirMap) delete kv.second; for (auto& kv : mNamed) delete kv.second; }

    // c = a + b on c.size() bits (operands sign- or zero-extended)
    static void add_build(BetaCircuit& cd, const BetaBundle& a, const BetaBundle& b, const BetaBundle& c, const BetaBundle& /*temps*/,
                          IntType it, Optimized op) {
        if (op == Optimized::Depth) prefixAdd(cd, a, b, c, it, c.size());
        else rippleAdd(cd, a, b, c, it);
    }
    // c[0] = bit `bitIdx` of a + b (or a - b)
    static void extractBit_build(BetaCircuit& cd, const BetaBundle& a, const BetaBundle& b, const BetaBundle& c, const BetaBundle& /*temps*/,
                                 u64 bitIdx, IntType it, AdderType at, Optimized) {
        if (at != AdderType::Addition) throw std::runtime_error("BetaLibrary shim: extractBit of a difference is not needed on the path " LOCATION);
        if (c.size() != 1) throw RTE_LOC;
        BetaBundle sum(bitIdx + 1);
        cd.addTempWireBundle(sum);
        prefixAdd(cd, a, b, sum, it, bitIdx + 1);
        cd.addCopy(sum[bitIdx], c[0]);
    }
    static void lessThan_build(BetaCircuit& cd, const BetaBundle& a, const BetaBundle& b, const BetaBundle& c, IntType it, Optimized) {
        if (a.size() != b.size() || c.size() != 1) throw RTE_LOC;
        const u64 n = a.size();
        // a < b  <=>  sign of the (n+1)-bit difference a - b = a + ~b + 1 (signed: sign-extended; unsigned: zero-extended)
        std::vector<BetaWire> g(n), p(n);
        for (u64 i = 0; i < n; ++i) {
            cd.addTempWire(g[i]); cd.addTempWire(p[i]);
            cd.addGate(b[i], a[i], GateType::na_And, g[i]);        // a & ~b
            cd.addGate(a[i], b[i], GateType::Nxor, p[i]);          // a ^ ~b
        }
        BetaWire g0; cd.addTempWire(g0);
        cd.addGate(g[0], p[0], GateType::Xor, g0);                 // carry-in 1 folded into bit 0
        g[0] = g0;
        std::vector<BetaWire> G = prefixAll(cd, g, p);
        if (it == IntType::TwosComplement) cd.addGate(p[n - 1], G[n - 1], GateType::Xor, c[0]);
        else cd.addInvert(G[n - 1], c[0]);                         // unsigned: borrow = NOT carry
    }
    // c = choice ? a : b   (choice has one wire)
    static void multiplex_build(BetaCircuit& cd, const BetaBundle& a, const BetaBundle& b, const BetaBundle& choice, const BetaBundle& c, const BetaBundle&) {
        for (u64 i = 0; i < c.size(); ++i) {
            BetaWire x, y; cd.addTempWire(x); cd.addTempWire(y);
            cd.addGate(a[i], b[i], GateType::Xor, x);
            cd.addGate(x, choice[0], GateType::And, y);
            cd.addGate(y, b[i], GateType::Xor, c[i]);
        }
    }

    BetaCircuit* int_int_bitwiseAnd(u64 a, u64 b, u64 c) { return bitwise("and", GateType::And, a, b, c); }
    BetaCircuit* int_int_bitwiseOr(u64 a, u64 b, u64 c) { return bitwise("or", GateType::Or, a, b, c); }
    BetaCircuit* int_int_bitwiseXor(u64 a, u64 b, u64 c) { return bitwise("xor", GateType::Xor, a, b, c); }
    BetaCircuit* int_int_add(u64 aBits, u64 bBits, u64 cBits, Optimized op = Optimized::Size) {
        return named("add" + std::to_string(aBits) + "_" + std::to_string(bBits) + "_" + std::to_string(cBits) + (op == Optimized::Depth ? "d" : "s"),
                     [&](BetaCircuit& cd) {
            BetaBundle a(aBits), b(bBits), c(cBits), t;
            cd.addInputBundle(a); cd.addInputBundle(b); cd.addOutputBundle(c);
            add_build(cd, a, b, c, t, IntType::TwosComplement, op);
        });
    }
    // c = a - b = a + ~b + 1: the same prefix network with generate = a & ~b, propagate = a ^ ~b and the carry-in folded
    // into bit 0 (aby3-Basic/BoolBasic.cpp:174)
    BetaCircuit* int_int_subtract(u64 aBits, u64 bBits, u64 cBits, Optimized = Optimized::Size) {
        return named("sub" + std::to_string(aBits) + "_" + std::to_string(bBits) + "_" + std::to_string(cBits), [&](BetaCircuit& cd) {
            BetaBundle a(aBits), b(bBits), c(cBits);
            cd.addInputBundle(a); cd.addInputBundle(b); cd.addOutputBundle(c);
            const u64 n = cBits;
            std::vector<BetaWire> g(n), p(n);
            for (u64 i = 0; i < n; ++i) {
                const BetaWire x = ext(cd, a, i, IntType::TwosComplement), y = ext(cd, b, i, IntType::TwosComplement);
                cd.addTempWire(g[i]); cd.addTempWire(p[i]);
                cd.addGate(y, x, GateType::na_And, g[i]);          // x & ~y
                cd.addGate(x, y, GateType::Nxor, p[i]);            // x ^ ~y
            }
            // carry-in 1: carry out of bit 0 = g0 | p0 = g0 ^ p0 (g0 & p0 = 0); sum bit 0 = p0 ^ 1
            std::vector<BetaWire> gg = g;
            BetaWire g0; cd.addTempWire(g0);
            cd.addGate(g[0], p[0], GateType::Xor, g0);
            gg[0] = g0;
            std::vector<BetaWire> G = prefixAll(cd, gg, p);
            cd.addInvert(p[0], c[0]);
            for (u64 i = 1; i < n; ++i) cd.addGate(p[i], G[i - 1], GateType::Xor, c[i]);
        });
    }
    // Operand order: the only thing in the reference tree that pins it is aby3-Basic's caller, which feeds the circuit
    // (input 0 = B, input 1 = A) for "A < B" (BoolBasic.cpp:31-32) and whose test expects bool_cipher_lt(Y, X) to reveal
    // x > y (aby3_tests/BoolTest.cpp:68,122,283) -- i.e. the circuit answers "input 1 < input 0".
    BetaCircuit* int_int_lt(u64 aBits, u64 bBits) {
        return named("lt" + std::to_string(aBits) + "_" + std::to_string(bBits), [&](BetaCircuit& cd) {
            BetaBundle a(aBits), b(bBits), c(1);
            cd.addInputBundle(a); cd.addInputBundle(b); cd.addOutputBundle(c);
            lessThan_build(cd, b, a, c, IntType::TwosComplement, Optimized::Depth);
        });
    }
    BetaCircuit* int_eq(u64 bits) {
        return named("eq" + std::to_string(bits), [&](BetaCircuit& cd) {
            BetaBundle a(bits), b(bits), c(1);
            cd.addInputBundle(a); cd.addInputBundle(b); cd.addOutputBundle(c);
            std::vector<BetaWire> e(bits);
            for (u64 i = 0; i < bits; ++i) { cd.addTempWire(e[i]); cd.addGate(a[i], b[i], GateType::Nxor, e[i]); }
            while (e.size() > 1) {
                std::vector<BetaWire> nx;
                for (u64 i = 0; i + 1 < e.size(); i += 2) { BetaWire w; cd.addTempWire(w); cd.addGate(e[i], e[i + 1], GateType::And, w); nx.push_back(w); }
                if (e.size() & 1) nx.push_back(e.back());
                e.swap(nx);
            }
            cd.addCopy(e[0], c[0]);
        });
    }
private:
    template <typename F>
    BetaCircuit* named(const std::string& key, F build) {
        auto it = mNamed.find(key);
        if (it != mNamed.end()) return it->second;
        BetaCircuit* cd = new BetaCircuit;
        build(*cd);
        cd->levelByAndDepth();
        mNamed[key] = cd;
        return cd;
    }
    BetaCircuit* bitwise(const char* name, GateType t, u64 aBits, u64 bBits, u64 cBits) {
        if (aBits != bBits || aBits != cBits) throw RTE_LOC;
        return named(std::string(name) + std::to_string(aBits), [&](BetaCircuit& cd) {
            BetaBundle a(aBits), b(bBits), c(cBits);
            cd.addInputBundle(a); cd.addInputBundle(b); cd.addOutputBundle(c);
            for (u64 i = 0; i < aBits; ++i) cd.addGate(a[i], b[i], t, c[i]);
        });
    }
    static BetaWire ext(BetaCircuit&, const BetaBundle& x, u64 i, IntType it) {
        if (i < x.size()) return x[i];
        if (it == IntType::TwosComplement) return x[x.size() - 1];
        throw std::runtime_error("BetaLibrary shim: zero extension of unsigned operands is not needed on the path " LOCATION);
    }
    static void rippleAdd(BetaCircuit& cd, const BetaBundle& a, const BetaBundle& b, const BetaBundle& c, IntType it) {
        const u64 n = c.size();
        BetaWire carry = BetaWire(-1);
        for (u64 i = 0; i < n; ++i) {
            const BetaWire x = ext(cd, a, i, it), y = ext(cd, b, i, it);
            BetaWire axb; cd.addTempWire(axb);
            cd.addGate(x, y, GateType::Xor, axb);
            if (i == 0) {
                cd.addCopy(axb, c[0]);
                if (n > 1) { cd.addTempWire(carry); cd.addGate(x, y, GateType::And, carry); }
            } else {
                cd.addGate(axb, carry, GateType::Xor, c[i]);
                if (i + 1 < n) {
                    BetaWire xc, yc, t, nc;
                    cd.addTempWire(xc); cd.addTempWire(yc); cd.addTempWire(t); cd.addTempWire(nc);
                    cd.addGate(x, carry, GateType::Xor, xc);
                    cd.addGate(y, carry, GateType::Xor, yc);
                    cd.addGate(xc, yc, GateType::And, t);
                    cd.addGate(t, carry, GateType::Xor, nc);
                    carry = nc;
                }
            }
        }
    }
    static std::vector<BetaWire> prefixAll(BetaCircuit& cd, std::vector<BetaWire> g, std::vector<BetaWire> p) {
        const u64 n = g.size();
        for (u64 d = 1; d < n; d <<= 1) {
            std::vector<BetaWire> ng = g, np = p;
            for (u64 i = d; i < n; ++i) {
                BetaWire t, G, P;
                cd.addTempWire(t); cd.addTempWire(G);
                cd.addGate(p[i], g[i - d], GateType::And, t);
                cd.addGate(g[i], t, GateType::Xor, G);
                ng[i] = G;
                cd.addTempWire(P);
                cd.addGate(p[i], p[i - d], GateType::And, P);
                np[i] = P;
            }
            g.swap(ng); p.swap(np);
        }
        return g;
    }
    static void prefixAdd(BetaCircuit& cd, const BetaBundle& a, const BetaBundle& b, const BetaBundle& c, IntType it, u64 n) {
        std::vector<BetaWire> g(n), p(n);
        for (u64 i = 0; i < n; ++i) {
            cd.addTempWire(g[i]); cd.addTempWire(p[i]);
            cd.addGate(ext(cd, a, i, it), ext(cd, b, i, it), GateType::And, g[i]);
            cd.addGate(ext(cd, a, i, it), ext(cd, b, i, it), GateType::Xor, p[i]);
        }
        std::vector<BetaWire> G = prefixAll(cd, g, p);
        cd.addCopy(p[0], c[0]);
        for (u64 i = 1; i < n; ++i) cd.addGate(p[i], G[i - 1], GateType::Xor, c[i]);
    }
};
}  // namespace osuCrypto
