"""Alias of the reference's top-level package for running its test files against shogidrl_b200 (tests only)."""
