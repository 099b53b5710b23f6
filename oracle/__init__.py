"""Test infrastructure only (see fs2_oracle.py).  Never imported by the product package."""
