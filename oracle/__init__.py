"""CPU oracle of the Neural MMO step -- TEST INFRASTRUCTURE, never imported by nmmo_b200."""
