"""Test infrastructure: CPU oracle for the CPG masked-conv hot path (see cpg_oracle.py)."""
