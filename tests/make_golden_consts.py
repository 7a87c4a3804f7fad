"""Constants shared by oracle/make_golden.py and the tests (kept import-light)."""
FIELD_TIMES = [0.0, 1.0 / 6.0, 1.0 / 3.0, 0.37, 0.5, 5.0 / 6.0, 1.0]
GAMMAS2 = [(0.0, 0.0), (0.2, 0.0), (0.2, 0.2), (1.0, 2.0), (-0.5, 0.3)]
GAMMAS3 = [(0.0, 0.0, 0.0), (0.2, 0.2, 0.2), (1.0, 2.0, 0.5)]
