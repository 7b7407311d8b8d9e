"""Mirror of the reference's ``sig`` package (src/sig/): perceptual hashes on the GPU."""
