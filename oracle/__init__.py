"""CPU oracle for the IAN hot path -- TEST INFRASTRUCTURE ONLY (parity unpinned, see ian_numpy.py).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference.
"""
