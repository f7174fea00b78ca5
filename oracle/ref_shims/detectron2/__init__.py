"""Minimal detectron2 stand-in used ONLY to import the read-only reference
(/root/reference) in the build container for oracle validation and golden-vector
generation.  Test infrastructure; never imported by the product path."""
