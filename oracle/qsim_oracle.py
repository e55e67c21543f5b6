"""CPU oracle for the statevector hot path of justinbrianhwang/Quantum-Simulator.

TEST INFRASTRUCTURE ONLY.  This module is a NumPy restatement of the reference
engine's algorithms (`quantum_sim/engine/*.py` under /root/reference).  It is
imported only by `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs, always as the *checker* (or the
timed CPU baseline) and never as the thing shipped: nothing under
`quantum-simulator_b200/` imports it, and the product path raises when the CUDA
library is missing instead of coming here.

Parity status: PINNED.  `tests/golden/make_golden.py` imports the real
reference in the build container and freezes its outputs into
`tests/golden/*.json|npz`; `tests/test_oracle_golden.py` checks every function
below against those fixtures (amplitudes to 1e-13, counts / branches /
syndromes bit-exact).

All citations are `file:line` relative to /root/reference.

Data model (plain Python, no reference classes):
  gate   = (name: str, targets: list[int], params: list[float], column: int)
  noise  = {"global": [(kind, p), ...], "gate": {gate_name: [(kind, p), ...]},
            "readout": (p01, p10) | None}
           kind in {"bit_flip", "phase_flip", "depolarizing", "amplitude_damping"}
  states = complex128[2**n], qubit 0 = most significant bit of the index
           (state_vector.py:87-88, :171-175)
"""

from __future__ import annotations

import math

import numpy as np

# --------------------------------------------------------------------------
# Gate tables (gates.py:37-125, gate_registry.py:34-148)
# --------------------------------------------------------------------------

_SQ2 = np.sqrt(2)


def _mat(rows):
    return np.array(rows, dtype=np.complex128)


def _perm_matrix(dim, swaps):
    m = np.eye(dim, dtype=np.complex128)
    for a, b in swaps:
        m[[a, b]] = m[[b, a]]
    return m


_FIXED = {
    "I": _mat([[1, 0], [0, 1]]),                       # gates.py:37
    "X": _mat([[0, 1], [1, 0]]),                       # gates.py:39-40
    "Y": _mat([[0, -1j], [1j, 0]]),                    # gates.py:42-43
    "Z": _mat([[1, 0], [0, -1]]),                      # gates.py:45-46
    "H": _mat([[1, 1], [1, -1]]) / _SQ2,               # gates.py:48-49
    "S": _mat([[1, 0], [0, 1j]]),                      # gates.py:51-52
    "S_DAG": _mat([[1, 0], [0, -1j]]),                 # gates.py:54-55
    "T": _mat([[1, 0], [0, np.exp(1j * np.pi / 4)]]),  # gates.py:57-58
    "T_DAG": _mat([[1, 0], [0, np.exp(-1j * np.pi / 4)]]),  # gates.py:60-61
    "CNOT": _perm_matrix(4, [(2, 3)]),                 # gates.py:99-103
    "CZ": np.diag([1, 1, 1, -1]).astype(np.complex128),  # gates.py:105
    "SWAP": _perm_matrix(4, [(1, 2)]),                 # gates.py:107-111
    "Toffoli": _perm_matrix(8, [(6, 7)]),              # gates.py:114-118
    "Fredkin": _perm_matrix(8, [(5, 6)]),              # gates.py:121-125
    "Measure": _mat([[1, 0], [0, 1]]),                 # gate_registry.py:136-139
    "Barrier": _mat([[1, 0], [0, 1]]),                 # gate_registry.py:142-145
}

SKIPPED = ("Measure", "Barrier")          # simulator.py:60-65
ARITY = {"CNOT": 2, "CZ": 2, "SWAP": 2, "Toffoli": 3, "Fredkin": 3}
NUM_PARAMS = {"Rx": 1, "Ry": 1, "Rz": 1, "Phase": 1, "U3": 3}


def gate_matrix(name, params=()):
    """Matrix of a registered gate (gates.py:66-94 for the parameterised ones)."""
    if name in _FIXED:
        return _FIXED[name]
    if name == "Rx":
        c, s = np.cos(params[0] / 2), np.sin(params[0] / 2)
        return _mat([[c, -1j * s], [-1j * s, c]])
    if name == "Ry":
        c, s = np.cos(params[0] / 2), np.sin(params[0] / 2)
        return _mat([[c, -s], [s, c]])
    if name == "Rz":
        t = params[0]
        return _mat([[np.exp(-1j * t / 2), 0], [0, np.exp(1j * t / 2)]])
    if name == "Phase":
        return _mat([[1, 0], [0, np.exp(1j * params[0])]])
    if name == "U3":
        th, ph, lam = params
        c, s = np.cos(th / 2), np.sin(th / 2)
        return _mat([[c, -np.exp(1j * lam) * s],
                     [np.exp(1j * ph) * s, np.exp(1j * (ph + lam)) * c]])
    raise KeyError(f"Gate '{name}' not found in registry")   # gate_registry.py:153-156


# --------------------------------------------------------------------------
# a1: StateVector.apply_gate (state_vector.py:41-74) incl. the position scramble
# --------------------------------------------------------------------------

def sigma(n, targets):
    """Axis permutation the reference leaves behind after apply_gate.

    state_vector.py:66-73 builds `dest_order` (the permutation that WOULD
    restore canonical order) and then transposes with `argsort(dest_order)`
    instead, so array axis i of the output holds textbook qubit
    sigma[i] = inv[inv[i]], inv = argsort(dest_order).
    """
    tset = set(targets)
    dest = [0] * n
    for i, q in enumerate(targets):
        dest[q] = i
    rest = [q for q in range(n) if q not in tset]
    for i, q in enumerate(rest):
        dest[q] = len(targets) + i
    inv = np.argsort(dest)
    return [int(inv[inv[i]]) for i in range(n)]


def apply_textbook(psi, n, matrix, targets):
    """U on `targets` (targets[0] = MSB of the matrix index), no scramble."""
    k = len(targets)
    t = psi.reshape([2] * n)
    t = np.moveaxis(t, targets, range(k)).reshape(2 ** k, -1)
    t = np.asarray(matrix, dtype=np.complex128).reshape(2 ** k, 2 ** k) @ t
    t = np.moveaxis(t.reshape([2] * n), range(k), targets)
    return np.ascontiguousarray(t).reshape(-1)


def apply_gate(psi, n, matrix, targets, layout="reference"):
    """state_vector.py:41-74.  Raises like :50-52 on a bad qubit index."""
    for q in targets:
        if q < 0 or q >= n:
            raise ValueError(f"Qubit index {q} out of range [0, {n-1}]")
    out = apply_textbook(psi, n, matrix, list(targets))
    if layout == "reference":
        s = sigma(n, list(targets))
        if s != list(range(n)):
            out = np.ascontiguousarray(out.reshape([2] * n).transpose(s)).reshape(-1)
    return out


def basis_state(initial_states):
    """state_vector.py:161-178."""
    n = len(initial_states)
    idx = 0
    for i, bit in enumerate(initial_states):
        if bit:
            idx |= 1 << (n - 1 - i)
    psi = np.zeros(2 ** n, dtype=np.complex128)
    psi[idx] = 1.0
    return psi


def probabilities(psi):
    """state_vector.py:36-39."""
    return np.abs(psi) ** 2


# --------------------------------------------------------------------------
# a3: gate ordering (circuit.py:69-79) and the run loop (simulator.py:35-114)
# --------------------------------------------------------------------------

def ordered_gates(gates):
    """Columns ascending, empty columns dropped, stable sort on targets[0]."""
    if not gates:
        return []
    cols = {}
    for g in gates:
        cols.setdefault(g[3], []).append(g)
    return [sorted(cols[c], key=lambda g: g[1][0]) for c in sorted(cols)]


# --------------------------------------------------------------------------
# a4: Kraus sets (noise.py:27-103) and the stochastic step (noise.py:224-260)
# --------------------------------------------------------------------------

def kraus_ops(kind, p):
    if not 0 <= p <= 1:
        raise ValueError("probability out of range")
    I, X, Y, Z = (_FIXED[k] for k in "IXYZ")
    if kind == "bit_flip":
        return [np.sqrt(1 - p) * I, np.sqrt(p) * X]
    if kind == "phase_flip":
        return [np.sqrt(1 - p) * I, np.sqrt(p) * Z]
    if kind == "depolarizing":
        return [np.sqrt(1 - p) * I, np.sqrt(p / 3) * X, np.sqrt(p / 3) * Y, np.sqrt(p / 3) * Z]
    if kind == "amplitude_damping":
        return [_mat([[1, 0], [0, np.sqrt(1 - p)]]), _mat([[0, np.sqrt(p)], [0, 0]])]
    raise KeyError(kind)


def choice_index(p, u):
    """numpy Generator.choice(k, p=p) given its single uniform draw `u`:
    cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, u, side='right')."""
    cdf = np.cumsum(np.asarray(p, dtype=np.float64))
    cdf /= cdf[-1]
    return int(np.searchsorted(cdf, u, side="right"))


def kraus_step(psi, n, ops, qubit, u, layout="reference"):
    """One (channel, qubit) draw: noise.py:241-260.  Returns (psi', branch)."""
    cands = [apply_gate(psi, n, K, [qubit], layout) for K in ops]
    probs = np.array([np.sum(np.abs(c) ** 2) for c in cands])
    tot = probs.sum()
    if tot > 1e-15:
        probs = probs / tot
    idx = choice_index(probs, u)
    out = cands[idx]
    nrm = np.sqrt(np.sum(np.abs(out) ** 2))
    if nrm > 1e-15:
        out = out / nrm
    return out, idx


def channels_for(noise, gate_name):
    """noise.py:217-219: global channels (insertion order) then gate-specific."""
    if noise is None:
        return []
    return list(noise.get("global", [])) + list(noise.get("gate", {}).get(gate_name, []))


def draw_count(n, gates, noise):
    """Uniform draws one trajectory consumes: one per (gate, channel, target<n)."""
    d = 0
    for col in ordered_gates(gates):
        for g in col:
            if g[0] in SKIPPED:
                continue
            d += len(channels_for(noise, g[0])) * sum(1 for q in g[1] if q < n)
    return d


def run_state(n, gates, initial=None, noise=None, draws=None, layout="reference",
              record_steps=False):
    """Gate loop of Simulator.run (simulator.py:53-71) without sampling.

    `draws` is the sequence of uniforms the noise model's private generator
    would return from `random()` (one per Kraus draw, noise.py:254).
    Returns (psi, steps, branches, has_measurement)."""
    psi = basis_state(initial if initial is not None else [0] * n)
    steps = [] if record_steps else None
    branches = []
    has_meas = False
    d = 0
    for col in ordered_gates(gates):
        for name, targets, params, _ in col:
            if name == "Measure":
                has_meas = True
                continue
            if name == "Barrier":
                continue
            psi = apply_gate(psi, n, gate_matrix(name, params), targets, layout)
            for kind, p in channels_for(noise, name):
                ops = kraus_ops(kind, p)
                for q in targets:
                    if q >= n:
                        continue
                    psi, b = kraus_step(psi, n, ops, q, draws[d], layout)
                    branches.append(b)
                    d += 1
        if record_steps:
            steps.append(psi.copy())
    return psi, steps, branches, has_meas


# --------------------------------------------------------------------------
# a7 / a8 / a9: readout + sampling (noise.py:120-175, measurement.py:38-129)
# --------------------------------------------------------------------------

def readout_distribution(probs, n, p01, p10):
    """ReadoutError.apply_to_distribution (noise.py:141-175)."""
    c = np.array([[1 - p01, p10], [p01, 1 - p10]])        # noise.py:120-126
    t = np.asarray(probs, dtype=np.float64).reshape([2] * n)
    for ax in range(n):
        t = np.moveaxis(np.tensordot(c, t, axes=([1], [ax])), 0, ax)
    out = np.ascontiguousarray(t).reshape(-1)
    tot = out.sum()
    if tot > 1e-15:
        out = out / tot
    return out


def readout_bitstring(bits, p01, p10, rng):
    """ReadoutError.apply_to_bitstring (noise.py:128-139)."""
    out = []
    for ch in bits:
        if ch == "0":
            out.append("1" if rng.random() < p01 else "0")
        else:
            out.append("0" if rng.random() < p10 else "1")
    return "".join(out)


def sample_counts(psi, n, shots, rng):
    """MeasurementEngine.sample (measurement.py:38-58)."""
    p = probabilities(psi)
    tot = p.sum()
    p = p / tot if tot > 1e-15 else np.ones_like(p) / len(p)
    c = rng.multinomial(shots, p)
    return {format(i, f"0{n}b"): int(v) for i, v in enumerate(c) if v > 0}


def rotate_basis(psi, n, basis, layout="reference"):
    """measurement.py:91-98: Y: S-dagger then H, X: H, on every qubit in order."""
    if basis == "Z":
        return psi
    out = psi
    for q in range(n):
        if basis == "Y":
            out = apply_gate(out, n, _FIXED["S_DAG"], [q], layout)
        out = apply_gate(out, n, _FIXED["H"], [q], layout)
    return out


def sample_with_basis(psi, n, shots, basis="Z", readout=None, readout_mode="shot",
                      rng=None, layout="reference"):
    """MeasurementEngine.sample_with_basis (measurement.py:60-129)."""
    rng = rng or np.random.default_rng()
    rot = rotate_basis(psi, n, basis, layout)
    if readout is not None and readout_mode == "distribution":
        p = probabilities(rot).copy()
        tot = p.sum()
        if tot > 1e-15:
            p /= tot
        noisy = readout_distribution(p, n, *readout)
        c = rng.multinomial(shots, noisy)
        return {format(i, f"0{n}b"): int(v) for i, v in enumerate(c) if v > 0}
    counts = sample_counts(rot, n, shots, rng)
    if readout is not None and readout_mode == "shot":
        noisy = {}
        for bits, cnt in counts.items():
            for _ in range(cnt):
                nb = readout_bitstring(bits, readout[0], readout[1], rng)
                noisy[nb] = noisy.get(nb, 0) + 1
        counts = noisy
    return counts


def run(n, gates, initial=None, noise=None, noise_seed=None, shots=1024, seed=None,
        basis="Z", record_steps=False, layout="reference"):
    """Simulator.run (simulator.py:35-91).  `noise_seed` = NoiseModel.set_seed."""
    rng = np.random.default_rng(seed)
    draws = None
    if noise is not None:
        draws = np.random.default_rng(noise_seed).random(draw_count(n, gates, noise))
    psi, steps, _, has_meas = run_state(n, gates, initial, noise, draws, layout, record_steps)
    counts = {}
    if has_meas or shots > 0:
        ro = noise.get("readout") if noise is not None else None
        counts = sample_with_basis(psi, n, shots, basis, ro, "shot", rng, layout)
    return psi, counts, steps


# --------------------------------------------------------------------------
# a5: run_with_noise (simulator.py:116-153) + measure_all (state_vector.py:107-119)
# --------------------------------------------------------------------------

def measure_all_index(psi, u):
    """state_vector.py:110-112 given the uniform `rng.choice` consumes."""
    p = probabilities(psi)
    p = p / p.sum()
    return choice_index(p, u)


def run_with_noise(n, gates, initial, noise, noise_seed, shots, seed, layout="reference"):
    """One noise generator shared by all shots (never reseeded), one
    `rng.choice` per shot on the measurement generator."""
    rng = np.random.default_rng(seed)
    nrng = np.random.default_rng(noise_seed)
    d = draw_count(n, gates, noise)
    counts = {}
    for _ in range(shots):
        psi, _, _, _ = run_state(n, gates, initial, noise, nrng.random(d), layout)
        idx = measure_all_index(psi, rng.random())
        key = format(idx, f"0{n}b")
        counts[key] = counts.get(key, 0) + 1
    return counts


# --------------------------------------------------------------------------
# a6: ensemble_density_matrix (simulator.py:155-199)
# --------------------------------------------------------------------------

def trial_seeds(seed, n_trials):
    """simulator.py:175-180 / qec.py:574-585 child-seed chain."""
    rng = np.random.default_rng(seed)
    return [int(rng.integers(0, 2 ** 63)) for _ in range(n_trials)]


def ensemble_states(n, gates, initial, noise, n_trials, seed, layout="reference"):
    d = draw_count(n, gates, noise)
    out = np.empty((n_trials, 2 ** n), dtype=np.complex128)
    for i, ts in enumerate(trial_seeds(seed, n_trials)):
        draws = np.random.default_rng(ts).random(d) if noise is not None else None
        out[i] = run_state(n, gates, initial, noise, draws, layout)[0]
    return out


def ensemble_density_matrix(n, gates, initial, noise, n_trials, seed, layout="reference"):
    """rho = (1/N) sum_i psi_i psi_i^dagger, no symmetrisation (simulator.py:195-198)."""
    psis = ensemble_states(n, gates, initial, noise, n_trials, seed, layout)
    return (psis.T @ psis.conj()) / n_trials


# --------------------------------------------------------------------------
# a10 / a11 / a12: reduced density matrices, entropies, MI, overlaps
# --------------------------------------------------------------------------

def partial_trace(psi, n, keep):
    """analysis.py:120-166 restated in O(2^n) memory: keep sorted, first kept
    qubit = MSB of the output index, rho = sum_env psi psi^*."""
    keep = sorted(keep)
    k = len(keep)
    m = np.moveaxis(psi.reshape([2] * n), keep, range(k)).reshape(2 ** k, -1)
    return m @ m.conj().T


def reduced_density_matrix_1q(psi, n, qubit):
    """state_vector.py:121-140."""
    t = psi.reshape(2 ** qubit, 2, 2 ** (n - qubit - 1))
    return np.einsum("aib,ajb->ij", t, t.conj())


def entropy_bits(rho):
    """analysis.py:99-104."""
    w = np.linalg.eigvalsh(rho)
    w = w[w > 1e-15]
    return float(-np.sum(w * np.log2(w)))


def entanglement_entropy(psi, n, qubits):
    """analysis.py:106-116."""
    return entropy_bits(partial_trace(psi, n, qubits))


def mutual_information(psi, n, a, b):
    """analysis.py:183-191."""
    return float(max(0.0, entanglement_entropy(psi, n, [a]) + entanglement_entropy(psi, n, [b])
                     - entanglement_entropy(psi, n, [a, b])))


def all_pairs_mi(psi, n):
    """Pair order of EntanglementEventDetector.process_step (analysis.py:331-333)."""
    return [mutual_information(psi, n, i, j) for i in range(n) for j in range(i + 1, n)]


def state_fidelity(psi, phi):
    """analysis.py:37-40."""
    return float(np.abs(np.vdot(psi, phi)) ** 2)


def expectation_value(psi, n, observable, targets, layout="reference"):
    """analysis.py:222-237: <psi| (O applied through apply_gate, scramble included)."""
    return complex(np.vdot(psi, apply_gate(psi, n, observable, targets, layout)))


def pauli_expectation(psi, n, pauli, qubit, layout="reference"):
    """analysis.py:239-248."""
    return float(np.real(expectation_value(psi, n, _FIXED[pauli.upper()], [qubit], layout)))


def purity_dm(rho):
    """analysis.py:176-179."""
    return float(np.real(np.trace(rho @ rho)))


# --------------------------------------------------------------------------
# a13: QEC cycle pieces (qec.py) -- Steane [[7,1,3]] and the two repetition codes
# --------------------------------------------------------------------------

_HAMMING_GEN = np.array([[1, 0, 0, 0, 1, 1, 0], [0, 1, 0, 0, 1, 0, 1],
                         [0, 0, 1, 0, 0, 1, 1], [0, 0, 0, 1, 1, 1, 1]])   # qec.py:364-369
STEANE_CHECKS = [[0, 2, 4, 6], [1, 2, 5, 6], [3, 4, 5, 6]]                 # qec.py:337-338


def steane_encode(logical):
    """qec.py:356-397: 8 codewords of the right weight parity, ancillas |0>."""
    n = 13
    words = [tuple(int(v) for v in (np.array([(i >> b) & 1 for b in range(4)]) @ _HAMMING_GEN) % 2)
             for i in range(16)]
    sel = [w for w in words if sum(w) % 2 == logical]
    psi = np.zeros(2 ** n, dtype=np.complex128)
    amp = 1.0 / np.sqrt(len(sel))
    for w in sel:
        idx = 0
        for q, bit in enumerate(w):
            if bit:
                idx |= 1 << (n - 1 - q)
        psi[idx] = amp
    return psi


def repetition_encode(logical, phase):
    """BitFlipCode.encode / PhaseFlipCode.encode (qec.py:185-195, :257-271)."""
    gates = []
    if logical == 1:
        gates.append(("X", [0], [], 0))
    gates += [("CNOT", [0, 1], [], 1), ("CNOT", [0, 2], [], 2)]
    if phase:
        gates += [("H", [q], [], 3) for q in range(3)]
    return run_state(5, gates)[0]


def z_parity_weights(psi, n, qubits):
    """(p_even, p_odd) of qec.py:466-484, vectorised."""
    idx = np.arange(2 ** n)
    par = np.zeros(2 ** n, dtype=np.int64)
    for q in qubits:
        par ^= (idx >> (n - 1 - q)) & 1
    p = probabilities(psi)
    return float(p[par == 0].sum()), float(p[par == 1].sum())


def z_parity(psi, n, qubits):
    """qec.py:486: ties go to 0."""
    e, o = z_parity_weights(psi, n, qubits)
    return 0 if e >= o else 1


def logical_z(psi, n, qubits):
    """QECCode.logical_z_expectation (qec.py:131-151)."""
    e, o = z_parity_weights(psi, n, qubits)
    return e - o


def qec_noise_paulis(noise_type, prob, n_data, rng):
    """QECSimulator._apply_noise (qec.py:669-693): one rng.random() per data
    qubit; returns the fired [(pauli, qubit)] list in application order."""
    fired = []
    for q in range(n_data):
        r = rng.random()
        if noise_type == "bit_flip":
            if r < prob:
                fired.append(("X", q))
        elif noise_type == "phase_flip":
            if r < prob:
                fired.append(("Z", q))
        elif noise_type == "depolarizing":
            if r < prob / 3:
                fired.append(("X", q))
            elif r < 2 * prob / 3:
                fired.append(("Y", q))
            elif r < prob:
                fired.append(("Z", q))
    return fired


_CODES = {
    "steane": dict(n=13, data=7, zl=list(range(7))),
    "bit_flip": dict(n=5, data=3, zl=[0, 1, 2]),
    "phase_flip": dict(n=5, data=3, zl=[0, 1, 2]),
}


def _encode(code, logical):
    if code == "steane":
        return steane_encode(logical)
    return repetition_encode(logical, phase=(code == "phase_flip"))


def _h_all(psi, n, count, layout):
    for q in range(count):
        psi = apply_gate(psi, n, _FIXED["H"], [q], layout)
    return psi


def extract_syndrome(code, psi, layout="reference"):
    """SteaneCode (qec.py:399-417), BitFlipCode (:197-207), PhaseFlipCode (:273-287)."""
    n = _CODES[code]["n"]
    if code == "steane":
        xs = [z_parity(psi, n, c) for c in STEANE_CHECKS]
        tmp = _h_all(psi, n, 7, layout)
        return xs + [z_parity(tmp, n, c) for c in STEANE_CHECKS]
    tmp = _h_all(psi, n, 3, layout) if code == "phase_flip" else psi
    return [z_parity(tmp, n, [0, 1]), z_parity(tmp, n, [1, 2])]


def decode_syndrome(code, s):
    """qec.py:419-439 (Steane), :209-219 / :289-299 (repetition codes)."""
    if code == "steane":
        out = []
        zi = s[3] + 2 * s[4] + 4 * s[5]
        if 0 < zi <= 7:
            out.append(("X", zi - 1))
        xi = s[0] + 2 * s[1] + 4 * s[2]
        if 0 < xi <= 7:
            out.append(("Z", xi - 1))
        return out
    g = "X" if code == "bit_flip" else "Z"
    table = {(1, 0): 0, (1, 1): 1, (0, 1): 2}
    key = (s[0], s[1])
    return [(g, table[key])] if key in table else []


def qec_cycle(code, logical, noise_type, prob, seed, layout="reference"):
    """QECSimulator.run_cycle (qec.py:497-549).  Returns a dict of the scalars."""
    spec = _CODES[code]
    n = spec["n"]
    rng = np.random.default_rng(seed)
    ideal = _encode(code, logical)
    noisy = ideal
    fired = qec_noise_paulis(noise_type, prob, spec["data"], rng)
    for pauli, q in fired:
        noisy = apply_gate(noisy, n, _FIXED[pauli], [q], layout)
    syn = extract_syndrome(code, noisy, layout)
    corr = decode_syndrome(code, syn)
    fixed = noisy
    for pauli, q in corr:
        if q < n:
            fixed = apply_gate(fixed, n, _FIXED[pauli], [q], layout)
    if code == "phase_flip":                       # qec.py:309-318
        z = logical_z(_h_all(fixed, n, 3, layout), n, spec["zl"])
    else:
        z = logical_z(fixed, n, spec["zl"])
    sign = 1.0 if logical == 0 else -1.0
    return dict(fired=fired, syndrome=syn, corrections=corr,
                fidelity_before=state_fidelity(ideal, noisy),
                fidelity_after=state_fidelity(ideal, fixed),
                z_exp=z, logical_error=bool(z * sign < 0))


def threshold_sweep(code, noise_probs, n_trials, noise_type, seed, layout="reference"):
    """QECSimulator.threshold_sweep (qec.py:551-622)."""
    rng = np.random.default_rng(seed)
    out = []
    for p in noise_probs:
        succ = 0
        fid = zf = 0.0
        zok = 0
        for trial in range(n_trials):
            ts = int(rng.integers(0, 2 ** 63))
            r = qec_cycle(code, trial % 2, noise_type, p, ts, layout)
            succ += r["fidelity_after"] > 0.5
            fid += r["fidelity_after"]
            zf += abs(r["z_exp"])
            zok += not r["logical_error"]
        out.append(dict(physical_rate=p, logical_rate=1.0 - succ / n_trials,
                        success_rate=succ / n_trials, avg_fidelity=fid / n_trials,
                        logical_z_fidelity=zf / n_trials,
                        decoder_success_rate=zok / n_trials,
                        projection_logical_rate=1.0 - fid / n_trials))
    return out


# --------------------------------------------------------------------------
# a14: parameter batches (optimizer.py:66-88)
# --------------------------------------------------------------------------

def param_slots(gates):
    """auto_detect order (optimizer.py:74-88): (gate_index, param_index) by
    gate insertion order."""
    return [(gi, pi) for gi, g in enumerate(gates)
            for pi in range(NUM_PARAMS.get(g[0], 0))]


def bind_values(gates, values):
    """optimizer.py:66-72."""
    out = [(g[0], list(g[1]), list(g[2]), g[3]) for g in gates]
    for v, (gi, pi) in zip(values, param_slots(gates)):
        out[gi][2][pi] = float(v)
    return out


def vqe_cost(psi, n, terms, layout="reference"):
    """CostFunction.vqe_hamiltonian (optimizer.py:133-165)."""
    tot = 0.0
    for coeff, label, qubits in terms:
        obs = _FIXED[label[0]]
        for ch in label[1:]:
            obs = np.kron(obs, _FIXED[ch])
        tot += coeff * float(np.real(expectation_value(psi, n, obs, list(qubits), layout)))
    return tot


# --------------------------------------------------------------------------
# Counter-based RNG used by the CUDA path's "philox" draw mode (ours, not the
# reference's): Philox4x32-10, key = seed, counter = (trajectory, draw block).
# Kept here so tests can feed the oracle the exact uniforms the kernel derives.
# --------------------------------------------------------------------------

def philox4x32_10(counter, key):
    m0, m1 = 0xD2511F53, 0xCD9E8D57
    w0, w1 = 0x9E3779B9, 0xBB67AE85
    c = [int(x) & 0xFFFFFFFF for x in counter]
    k = [int(x) & 0xFFFFFFFF for x in key]
    for _ in range(10):
        p0 = m0 * c[0]
        p1 = m1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF,
             ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + w0) & 0xFFFFFFFF, (k[1] + w1) & 0xFFFFFFFF]
    return c


def philox_uniforms(seed, trajectory, count):
    """uniform[d] for d < count: block d//2, words (2*(d%2), 2*(d%2)+1),
    u = ((hi >> 5) * 2^26 + (lo >> 6)) * 2^-53."""
    out = np.empty(count, dtype=np.float64)
    for blk in range((count + 1) // 2):
        r = philox4x32_10([trajectory & 0xFFFFFFFF, (trajectory >> 32) & 0xFFFFFFFF, blk, 0],
                          [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF])
        for h in range(2):
            d = 2 * blk + h
            if d < count:
                out[d] = ((r[2 * h + 1] >> 5) * 67108864.0 + (r[2 * h] >> 6)) * (1.0 / 9007199254740992.0)
    return out
