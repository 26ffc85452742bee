"""CPU oracle of the MaxK-GNN aggregation hot path: test infrastructure, not product code.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.
"""
