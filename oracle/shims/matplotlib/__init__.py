"""Empty stand-in so `import matplotlib.pyplot` in the reference succeeds (test infrastructure only)."""
