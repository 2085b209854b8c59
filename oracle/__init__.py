"""CPU oracle of the LLGS hot path — TEST INFRASTRUCTURE ONLY (see the header of each module).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
