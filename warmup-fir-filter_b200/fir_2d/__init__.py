"""2-D counterpart of ``fir_1d`` (the reference's ``fir_2d/`` is an empty stub;
the specification is in DESIGN.md)."""
