"""fdm_b200: host side of the B200-native LG-LDM sampling path (ctypes binding + sampling engine)."""
