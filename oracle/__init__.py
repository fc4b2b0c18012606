"""oracle/ — TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import anything from here; the product (s2geometry_b200/) never does."""
