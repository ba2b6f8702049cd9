"""Baseline harnesses: the UNMODIFIED reference (vendored copy under baseline/_ref, git-ignored) driven on synthetic
data, used by ``bench.py --impl reference``.  Nothing in here is part of the product path."""
