"""`NewFunctions/Gotennet_morethaninspired` (HTR edge update + GATA value activation; used by
equiformerv2_MatPES_GATAV2.py)."""
