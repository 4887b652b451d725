"""CPU oracles (test infrastructure).  See the header of each module.

Nothing under nylon_amt_b200/ imports this package; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs do.
"""
