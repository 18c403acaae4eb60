"""Import-time stand-in for TensorFlow so the reference's numpy-only members
(loaders, sampler, .dat codec) can be imported in the build container to
generate golden fixtures.  No TensorFlow behaviour is emulated."""
