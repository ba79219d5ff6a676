run pipe A=1
run nopipe --args "--no-pipeline" A=1
run pipe10 --args "--steps 20" A=1
