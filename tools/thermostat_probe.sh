#!/bin/bash
# throughput of the host-driven integrators through lmp_b200 (4 M-atom LJ melt, 20 + 100 steps):
# fix nvt, fix npt, fix nve + fix langevin; $1 = extra package options (e.g. "lazy no")
exe=lammps_b200/lammps_pkg/lmp_b200
for fix in "fix 1 all nvt temp 1.44 1.44 0.5" "fix 1 all npt temp 1.44 1.44 0.5 iso 5.0 5.0 5.0" "fix 1 all nve
fix 2 all langevin 1.44 1.44 1.0 48279" "fix 1 all nve"; do
  printf 'units lj\nlattice fcc 0.8442\nregion box block 0 100 0 100 0 100\ncreate_box 1 box\ncreate_atoms 1 box\nmass 1 1.0\nvelocity all create 1.44 87287 loop geom\npair_style lj/cut 2.5\npair_coeff 1 1 1.0 1.0 2.5\nneighbor 0.3 bin\nneigh_modify delay 0 every 20 check no\n%s\nthermo 100\nrun 20\nrun 100\n' "$fix" > /tmp/in.thermo
  out=$($exe -sf b200 -pk b200 $1 -echo none -in /tmp/in.thermo 2>&1 | grep "Loop time" | tail -1)
  echo "$(echo $fix | tr '\n' ';') [$1] -> $out"
done
