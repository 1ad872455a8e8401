bash tools/r2_test.sh $1
bash tools/r2_prof_others.sh $1
