"""iqtree2_b200 — B200-native drop-in for IQ-TREE 2's PhyloTree likelihood path."""
__version__ = "0.1.0"
