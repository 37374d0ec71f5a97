"""TESTS ONLY -- empty stand-in; nothing on the particle path touches yt."""
