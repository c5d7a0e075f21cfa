"""Plain-PyTorch models for the REFERENCE arm of bench.py (no adaptdl_b200
imports): used where the reference's own example model cannot be imported on
this stack."""
