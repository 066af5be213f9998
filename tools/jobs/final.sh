# the verification pass behind the round's closing numbers (profiles/r02j_*)
bash tools/jobs/check.sh smoke tests bench:cfg3 ref bench:cfg2
bash tools/jobs/profile.sh cfg3 launches
bash tools/jobs/check.sh stress:cfg3:60:2
