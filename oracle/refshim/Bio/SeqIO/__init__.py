"""Minimal stand-in for Biopython (TEST INFRASTRUCTURE ONLY): the reference imports two parsers."""
