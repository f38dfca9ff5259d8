/* TEST INFRASTRUCTURE stub: nothing of sam.h is used by call_vars()/report_var() */
