"""Gate classes of the B200 build (same module / class names as the reference's MPDOSimulator.QuantumGates)."""
