"""Native extension packages of the B200 backend (mirrors ``src/extensions_ref``)."""
