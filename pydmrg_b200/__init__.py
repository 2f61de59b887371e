"""pydmrg_b200 -- B200-native (sm_100a) implementation of PyDMRG's DMRG sweep hot path."""
__version__ = "0.1.0"
