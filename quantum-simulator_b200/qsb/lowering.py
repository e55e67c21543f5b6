"""Circuit (+ noise model) -> device program.

Restates the control flow of the reference's run loop -- `Simulator.run` (simulator.py:53-71),
`NoiseModel.apply` (noise.py:212-222) -- as a lowering pass: gate order, Measure/Barrier skipping,
"noise after every executed gate: global channels then gate-specific, one draw per (channel, target)".
"""

from __future__ import annotations

from .compiler import Lowering, _PARAM_COUNT

SKIP_TYPES = ("measurement", "barrier")


def lower_circuit(n, columns, registry, channels_of=None, *, record_steps=False, param_offsets=None,
                  layout="reference", local_bits=None, extra=None, stream=False, max_local_bits=13):
    """Lower ordered gate columns.

    columns      : circuit.get_ordered_gates() -- lists of objects with gate_name / target_qubits / params
    registry     : GateRegistry (get(name) -> GateDefinition; KeyError for unknown names)
    channels_of  : name -> [(kind, p, kraus_ops | None), ...] in application order, or None (noiseless)
    param_offsets: {id(gate): offset} -- those Rx/Ry/Rz/Phase/U3 gates read their angles from the
                   per-state parameter row instead of gate.params (config 2's parameter batches)
    extra        : callable(Lowering) appended after the circuit (basis rotations, observables)

    Returns (Program, has_measurement)."""
    lw = Lowering(n, layout=layout)
    has_meas = False
    for col in columns:
        for g in col:
            gd = registry.get(g.gate_name)
            gtype = gd.gate_type.value
            if gtype == "measurement":
                has_meas = True
                continue
            if gtype == "barrier":
                continue
            builtin = registry.is_builtin(g.gate_name) if hasattr(registry, "is_builtin") else True
            if param_offsets is not None and id(g) in param_offsets and g.gate_name in _PARAM_COUNT:
                if not builtin:
                    raise NotImplementedError(f"gate '{g.gate_name}' was re-registered: its angles cannot be bound on "
                                              "the device (only the built-in Rx/Ry/Rz/Phase/U3 formulas are)")
                lw.param_gate(g.gate_name, g.target_qubits, param_offsets[id(g)])
            else:
                lw.gate(g.gate_name, g.target_qubits, g.params, gd.matrix_func, builtin=builtin)
            if channels_of is not None:
                for kind, p, kraus_ops in channels_of(g.gate_name):
                    for q in g.target_qubits:
                        lw.kraus(kind, p, q, kraus_ops)
        if record_steps:
            lw.snapshot()
    if extra is not None:
        extra(lw)
    if stream:
        return lw.finish_stream(local_bits), has_meas
    return lw.finish(local_bits, max_local_bits), has_meas
