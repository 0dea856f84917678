"""tennis_b200 — B200-native (sm_100a) hot path of HaydenFaulkner/Tennis behind the reference's operator API."""
__version__ = "0.1.0"
