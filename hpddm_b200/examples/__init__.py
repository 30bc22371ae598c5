"""Driver-side utilities mirroring the reference's examples/ (generators, drivers)."""
