"""CPU oracle for the gaast phase-4 hot path.  TEST INFRASTRUCTURE ONLY: the
product package (`gaast_b200`) must never import anything from here."""
