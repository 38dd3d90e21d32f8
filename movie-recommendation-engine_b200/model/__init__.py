"""Drop-in mirrors of the reference's ``model`` package (inference hot path only)."""
