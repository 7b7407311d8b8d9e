"""Mirror of the reference's ``ui`` package limited to the duplicate-refinement helpers (no Qt)."""
