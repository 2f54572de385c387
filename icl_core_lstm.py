#!/usr/bin/env python
"""Drop-in for the reference's icl_core_lstm.py (same flags and outputs; see imagecaptionlearn_py_b200/drivers.py)."""
from imagecaptionlearn_py_b200.drivers import main_core

if __name__ == "__main__":
    main_core()
