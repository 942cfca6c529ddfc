"""CPU oracle -- test infrastructure only; never imported by sphy_b200."""
