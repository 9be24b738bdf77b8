"""TEST INFRASTRUCTURE ONLY - CPU restatements of the reference hot path used as the parity
checker.  Never imported by dpf_nets_b200 (see oracle/README.md)."""
