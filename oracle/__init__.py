"""CPU oracle package — TEST INFRASTRUCTURE ONLY (see oracle/odpd_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
