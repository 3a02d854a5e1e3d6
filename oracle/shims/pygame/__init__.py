"""Empty stand-in: the reference only touches pygame when params.render is True."""
