"""Drop-in mirrors of the reference's ``utils`` package (hot-path modules only)."""
