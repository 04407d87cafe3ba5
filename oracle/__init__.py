"""CPU oracle of the surfel render path — TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package (the product path under materialrefgs_b200/ never does).
"""
