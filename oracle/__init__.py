"""CPU oracle for the IAN hot path -- TEST INFRASTRUCTURE ONLY (parity pinned to the executed reference, see ian_numpy.py;
`refshim/` holds the numpy stand-ins for Theano/Lasagne that make executing it possible).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference.
"""
