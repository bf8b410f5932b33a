"""CPU oracle for the wflow_sbm hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  wflow_b200/ (the product) never does.
"""
