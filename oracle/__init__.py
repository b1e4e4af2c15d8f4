"""CPU oracle of the sim.py hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import anything from this package.  The product
(meshless_inflatable_softbody_b200) never does.  PARITY UNPINNED by the
reference: see mis_oracle.c.
"""
