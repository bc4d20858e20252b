"""Sections of bench.py (measurement infrastructure, NOT part of the cherryml_b200 product):
each function times a piece of the product on synthetic data and, on rank 0 at N = 1, the CPU
reference (oracle/_ref binaries) or oracle port beside it.  Like tests/ and bench.py itself, these
modules may import oracle/; nothing under cherryml_b200/ does."""
