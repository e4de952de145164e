"""Generated from csrc/qm_core.h (kinematics workspace offsets); test helper."""
KW = {'R': 0, 'P': 456, 'AX': 528, 'BODY': 216, 'COMP': 600, 'ACM': 840, 'SV': 984, 'V': 1128, 'HB': 1668, 'FPOS': 1272, 'FVEL': 1284, 'EEP': 1296, 'EER': 1299, 'COM': 1308, 'ABINV': 1311, 'VEL': 1347, 'RHS': 1371, 'VSIZE': 1380, 'FJ': 1380, 'EEJ': 1668, 'DH': 0, 'DFV': 144, 'F': 0, 'SIZE': 1812}
