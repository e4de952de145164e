"""Generated from csrc/qm_core.h (kinematics workspace offsets); test helper."""
KW = {'R': 0, 'P': 216, 'AX': 288, 'BODY': 360, 'COMP': 600, 'ACM': 840, 'SV': 984, 'V': 1128, 'HB': 1272, 'DH': 1416, 'FPOS': 1560, 'FVEL': 1572, 'FJ': 1584, 'DFV': 1872, 'EEP': 2160, 'EER': 2163, 'EEJ': 2172, 'COM': 2316, 'ABINV': 2319, 'VEL': 2355, 'RHS': 2379, 'F': 2385, 'SIZE': 2532.0}
