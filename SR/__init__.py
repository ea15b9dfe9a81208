"""Reference import path `SR.*` (train.py:14, mymodels.py:10-13) re-exported from the B200 package."""
