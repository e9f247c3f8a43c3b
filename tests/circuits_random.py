"""Random Boolean circuits in the flat layout the engines take (harness.library_circuit / orc_circuit): every gate
type the sh3 binary engine supports (Sh3BinaryEvaluator.cpp:700-1065), random fan-in from earlier wires, levelised
by AND depth with the linear gates of a level first (the facade's BetaCircuit::levelByAndDepth order), random
inverted outputs.  Used to compare the oracle with the reference's own evaluator and the device with the oracle."""
import numpy as np

XOR, AND, NOR, OR, NXOR, COPY, NA_AND = 6, 8, 1, 14, 9, 10, 4
LINEAR = {XOR, NXOR, COPY}
TYPES = [XOR, AND, NOR, OR, NXOR, COPY, NA_AND]


def random_circuit(seed, in_bits=(13, 7), n_gates=120, out_bits=(9, 3)):
    rng = np.random.default_rng(seed)
    wires = sum(in_bits)
    input_first, acc = [], 0
    for b in in_bits:
        input_first.append(acc)
        acc += b
    gates, ready = [], [0] * wires              # ready[w] = first level at which wire w is usable
    lvl = []
    for _ in range(n_gates):
        t = int(rng.choice(TYPES))
        a = int(rng.integers(0, wires))
        b = a
        if t != COPY:
            while b == a:
                b = int(rng.integers(0, wires))
        out = wires
        wires += 1
        l = max(ready[a], ready[b])
        gates.append((a, b, out, t))
        lvl.append(l)
        ready.append(l if t in LINEAR else l + 1)
    order = sorted(range(n_gates), key=lambda g: (lvl[g], 0 if gates[g][3] in LINEAR else 1, g))
    flat = np.array([gates[g] for g in order], dtype=np.uint32).reshape(-1)
    n_levels = max(lvl) + 1
    level_gates = np.zeros(n_levels, dtype=np.uint32)
    for g in range(n_gates):
        level_gates[lvl[g]] += 1
    output_wires, output_off, off = [], [], 0
    for b in out_bits:
        output_off.append(off)
        output_wires += [int(x) for x in rng.integers(0, wires, b)]
        off += b
    inv_wire = rng.integers(0, 2, wires)
    return {
        "wire_count": wires, "nonlinear": sum(1 for g in gates if g[3] not in LINEAR),
        "gates": flat, "level_gates": level_gates,
        "input_first": np.array(input_first, dtype=np.uint32), "input_bits": np.array(in_bits, dtype=np.uint32),
        "output_off": np.array(output_off, dtype=np.uint32), "output_bits": np.array(out_bits, dtype=np.uint32),
        "output_wires": np.array(output_wires, dtype=np.uint32),
        # inversion is a property of the WIRE (BetaWireFlag::InvWire), so every position reading a wire agrees
        "output_invert": np.array([inv_wire[w] for w in output_wires], dtype=np.uint8),
    }


def plain_eval(cir, inputs):
    """Plaintext evaluation: inputs = list of uint64 arrays [width] (one word per bundle, <= 64 bits each)."""
    width = len(inputs[0])
    vals = np.zeros((cir["wire_count"], width), dtype=np.uint8)
    for k, first in enumerate(cir["input_first"]):
        for j in range(int(cir["input_bits"][k])):
            vals[first + j] = (inputs[k] >> np.uint64(j)) & np.uint64(1)
    g = cir["gates"].reshape(-1, 4)
    for a, b, out, t in g:
        x, y = vals[a], vals[b]
        if t == XOR: v = x ^ y
        elif t == AND: v = x & y
        elif t == NOR: v = 1 - (x | y)
        elif t == OR: v = x | y
        elif t == NXOR: v = 1 - (x ^ y)
        elif t == COPY: v = x
        elif t == NA_AND: v = (1 - x) & y
        else: raise KeyError(t)
        vals[out] = v
    outs = []
    for k, off in enumerate(cir["output_off"]):
        word = np.zeros(width, dtype=np.uint64)
        for j in range(int(cir["output_bits"][k])):
            w = cir["output_wires"][off + j]
            bit = vals[w] ^ cir["output_invert"][off + j]
            word |= bit.astype(np.uint64) << np.uint64(j)
        outs.append(word)
    return outs
