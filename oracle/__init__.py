"""CPU oracle for the CamKifu stone-detection hot path — TEST INFRASTRUCTURE ONLY (see oracle/ck_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
"""
