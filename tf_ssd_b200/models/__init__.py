"""Mirror of the reference's ``models`` package."""
