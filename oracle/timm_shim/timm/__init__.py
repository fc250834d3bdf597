"""TEST INFRASTRUCTURE ONLY.  Minimal stand-in for the un-vendored dependency ``timm==0.9.2``
(reference requirements.yml:21) so that /root/reference/models/*.py can be imported in the build
container by oracle/make_golden.py.  It restates the published behaviour of the five classes the
reference uses (PatchEmbed, Attention, Block, Mlp, DropPath) with timm's sub-module names, so the
reference's ``state_dict`` key set is reproduced exactly.  Never imported by the product."""
__version__ = "0.9.2-shim"
