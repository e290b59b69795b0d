"""magphase_b200: B200-native MagPhase analysis/synthesis hot path (CUDA sm_100a behind a ctypes C ABI).

    import magphase_b200.magphase as mp      # same function names as the reference's src/magphase.py
"""
__version__ = '0.1'
