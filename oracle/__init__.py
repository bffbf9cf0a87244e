"""Test infrastructure only (CPU oracle for the TT hot path). See oracle/tt_oracle.py."""
