"""drop-in package: see ../README.md"""
