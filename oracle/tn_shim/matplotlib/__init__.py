"""Stub so that the reference's Tools.py (plotting helpers, out of scope) can be imported offline."""
rcParams = {}
