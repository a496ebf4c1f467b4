"""Recorded-data handling around the realtime reconstructor (host-side mirror of CEIT/datahandler)."""
