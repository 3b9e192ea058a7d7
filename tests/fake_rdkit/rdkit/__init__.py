"""Stand-in for the subset of RDKit's Python API that nvmolkit_b200/rdkit_adapter.py calls - TEST INFRASTRUCTURE ONLY.
Molecules wrap the synthetic pseudo-molecule dicts of nvmolkit_b200.synthetic (graph, coordinates, bounds, MMFF / UFF
parameter tables), so the adapters can be EXECUTED without RDKit and their output compared with the tables the generator
wrote directly. It pins index conventions and plumbing, not chemistry."""
