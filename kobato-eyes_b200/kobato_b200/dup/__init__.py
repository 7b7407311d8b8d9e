"""Mirror of the reference's ``dup`` package (src/dup/): candidate search and verification on the GPU."""
