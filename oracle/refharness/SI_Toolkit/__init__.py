"""TEST-ONLY shim package standing in for the un-vendored SI_Toolkit (see computation_library.py)."""
