"""The reference's global backend switch (globalvars.py:5).  This package has
only the device path, so the flag is informational."""
cuda = True
