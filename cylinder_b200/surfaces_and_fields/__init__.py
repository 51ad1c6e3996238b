"""Same module names as the reference's `surfaces_and_fields` package, so that
`import surfaces_and_fields.On_lattice_simple as lattice` becomes
`import cylinder_b200.surfaces_and_fields.On_lattice_simple as lattice`."""
