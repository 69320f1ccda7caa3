"""TEST INFRASTRUCTURE: CPU restatement of the reference's algorithm (ctmr_oracle.c, ctmr_oracle_frontend.c) and its ctypes
wrapper (oracle.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this."""
