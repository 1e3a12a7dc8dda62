"""CPU oracle (test infrastructure): see py_oracle.py and q1_port.c.  Never imported by the product."""
