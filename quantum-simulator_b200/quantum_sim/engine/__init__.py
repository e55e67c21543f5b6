"""Quantum simulation engine -- the reference's engine API on a B200 statevector backend."""
