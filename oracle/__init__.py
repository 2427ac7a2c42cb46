"""TEST INFRASTRUCTURE ONLY — CPU oracle of the Hair-GS render path and the recipe that builds the unmodified
reference as a parity oracle.  Nothing under hair-gs_b200/ may import this package (tests/test_host_cpu.py checks)."""
