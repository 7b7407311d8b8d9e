"""TEST INFRASTRUCTURE ONLY — a stand-in for the `skimage` package, which is absent from this image and cannot be
installed (no network).  It exists so that the LIVE reference's ``dup.refine`` (src/dup/refine.py:12 imports
``skimage.metrics.structural_similarity``) and ``dup.cluster`` import unmodified and their own code paths and tests run
here.  The one function behind it is oracle.ref_py's restatement of scikit-image 0.25.2's algorithm; the shim does not
make that restatement any more "pinned" against real scikit-image than it was."""
