"""Synthetic workloads shared by both benchmark arms and the tests (neutral: imports neither the product nor the oracle)."""
