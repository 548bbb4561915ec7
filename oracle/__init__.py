"""CPU oracle for the CTC path — test infrastructure only (see ctc_oracle.py header)."""
