"""Reference-side glue: read OpenMM objects (or anything with their methods) into the engine's parameter records."""
