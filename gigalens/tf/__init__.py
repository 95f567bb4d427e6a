"""``src/gigalens/tf``: the TensorFlow substrate's module paths, served by the CUDA implementation."""
