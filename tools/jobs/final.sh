bash tools/jobs/check.sh smoke tests bench:cfg3 ref bench:cfg2 sanitize
bash tools/jobs/profile.sh cfg3 launches metrics:edge_forward_tc2
bash tools/jobs/profile.sh mid full:edge_forward_tc2
bash tools/jobs/check.sh stress:cfg3:80:4
