"""CPU oracle for the replica-exchange hot path -- TEST INFRASTRUCTURE ONLY (see rx_oracle.c)."""
