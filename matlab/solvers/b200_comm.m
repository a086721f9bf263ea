function [rank, world, lo, hi] = b200_comm(h, m)
% B200_COMM  Row-sharded runs from MATLAB / Octave: ONE interpreter process per GPU (the engine's replacement for
% the PCT worker pool, admm.m:343-408, unwrappedadmm.m:45-74).  Each process is started with
%     ADMM_B200_RANK = 0..N-1,  ADMM_B200_WORLD = N,  ADMM_B200_ID_FILE = a path every process can read
% (and options.device = its GPU).  Rank 0 writes the 128-byte communicator id to the file, the others wait for it,
% everybody calls admm_b200_mex('comm_init', ...) once per handle.  Returns this rank's rows [lo, hi] of an m-row
% matrix by the reference's balancing rule (errorcheck.m:249-259).  Without the environment: rank 0 of 1, all rows.
%   Every process must draw the SAME random init (call rng(seed) identically before unwrappedadmm's rand calls):
%   a rank keeps its own rows of ONE global z0 / u0.  results.zopt / uopt hold THIS rank's rows.
% UNTESTED HERE (no MATLAB / Octave in the image); tested twin: admm_project_b200/parallel.py (attach_comm, row_range).
persistent attached
rank = str2double(getenv('ADMM_B200_RANK')); world = str2double(getenv('ADMM_B200_WORLD'));
if isnan(rank) || isnan(world) || world <= 1
    rank = 0; world = 1; lo = 1; hi = m; return;
end
if isempty(attached) || attached ~= h
    idfile = getenv('ADMM_B200_ID_FILE');
    if isempty(idfile), error('admm_b200: ADMM_B200_ID_FILE is not set (needed to share the communicator id).'); end
    if rank == 0
        id = admm_b200_mex('unique_id');
        f = fopen([idfile '.tmp'], 'w'); fwrite(f, id, 'uint8'); fclose(f);
        movefile([idfile '.tmp'], idfile);                      % appears atomically for the readers
    else
        t = tic;
        while ~exist(idfile, 'file')
            if toc(t) > 120, error('admm_b200: rank 0 did not publish the communicator id within 120 s.'); end
            pause(0.05);
        end
        f = fopen(idfile, 'r'); id = uint8(fread(f, 128, 'uint8'))'; fclose(f);
    end
    admm_b200_mex('comm_init', h, rank, world, id);
    attached = h;
end
sl = admm_b200_mex('slicemaker', m, world);
lo = sum(sl(1:rank)) + 1; hi = lo + sl(rank + 1) - 1;
end
