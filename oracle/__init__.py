"""ORACLE — test infrastructure only.

CPU restatement of google/forma's CPU path (see oracle/README.md). Only
tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
