"""Mirror of the reference's ``utils`` Cython modules on the hot path."""
