"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/horae_oracle.c header).  Never imported by horaedb_b200."""
