"""Mirror of the hot-path modules of the reference's ``core`` package (src/core/fastsig.py, signature.py)."""
