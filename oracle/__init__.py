"""TEST INFRASTRUCTURE ONLY: CPU oracle for the reference hot path (see yask_oracle.c).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
