"""TESTS ONLY -- empty stand-in so that ``ics.py`` imports; YAML round trips are out of scope."""
