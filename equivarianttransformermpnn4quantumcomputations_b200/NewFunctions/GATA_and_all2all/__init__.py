"""`NewFunctions/GATA_and_all2all`: global (all-to-all) node attention used by BASELINE config 5."""
