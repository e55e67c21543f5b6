"""Drop-in `quantum_sim` package: the reference's engine API backed by libqsb.so (B200, sm_100a).

Only `quantum_sim.engine` is provided here -- the statevector hot path.  GUI, controller,
bridge and core of the reference are out of scope and keep working on top of it unchanged.
"""
