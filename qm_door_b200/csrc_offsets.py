"""Generated from csrc/qm_core.h (kinematics workspace offsets); test helper."""
KW = {'R': 0, 'P': 456, 'AX': 528, 'BODY': 216, 'COMP': 600, 'ACM': 840, 'SV': 984, 'V': 1128, 'HB': 1272, 'FPOS': 1416, 'FVEL': 1428, 'EEP': 1440, 'EER': 1443, 'COM': 1452, 'ABINV': 1455, 'VEL': 1491, 'RHS': 1515, 'VSIZE': 1524, 'FJ': 1524, 'EEJ': 1812, 'DH': 0, 'DFV': 144, 'F': 0, 'SIZE': 1956}
