"""hoigen_b200 — B200-native (sm_100a) HOI scoring forward behind HOIGen's Python surface."""
__version__ = "0.1.0"
