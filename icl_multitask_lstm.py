#!/usr/bin/env python
"""Drop-in for the reference's icl_multitask_lstm.py (same flags and outputs; see imagecaptionlearn_py_b200/drivers.py)."""
from imagecaptionlearn_py_b200.drivers import main_multitask

if __name__ == "__main__":
    main_multitask()
