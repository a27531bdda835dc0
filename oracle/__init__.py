"""CPU oracle for the cyTVDN hot path -- test infrastructure only (see tv_oracle.py)."""
