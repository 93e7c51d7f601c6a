"""CPU oracle for the b2seg hot path — test infrastructure only (see keras_ref.py header)."""
